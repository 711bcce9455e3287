"""Shared test helpers: golden-fixture loading and the synthetic-input generators of SURVEY.md section 8(d)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class NbitCase:
    def __init__(self, z, row):
        name, w_bit, dt, asym, act, group, K, N = row.split(",")
        self.name, self.dt = name, dt
        self.w_bit, self.asym, self.act_order = int(w_bit), bool(int(asym)), bool(int(act))
        self.group, self.K, self.N = int(group), int(K), int(N)
        for key in ("qweight", "scales", "zeros", "g_idx", "x", "dy", "W", "y", "dx", "Wp", "packed",
                    "roundtrip_equal"):
            setattr(self, key, z[f"{name}_{key}"])

    def f(self, key):
        """float32 values of a 16-bit-stored array."""
        from oracle import nbit
        a = getattr(self, key)
        if self.dt == "f32" or a.dtype != np.uint16:
            return np.asarray(a, dtype=np.float32)
        return nbit.from_bits16(a, self.dt)

    @property
    def id(self):
        return f"{self.name}-b{self.w_bit}-{self.dt}-{'asym' if self.asym else 'sym'}" \
               f"{'-act' if self.act_order else ''}"


def load_nbit_cases():
    z = np.load(os.path.join(GOLD, "nbit_cases.npz"))
    return [NbitCase(z, str(r)) for r in z["cases"]]


# ---------------------------------------------------------------------------------------------------------------
# synthetic inputs of SURVEY.md section 8(d) (torch, device-agnostic) + torch<->numpy bridges for the oracle
# ---------------------------------------------------------------------------------------------------------------
def torch_dt(dt):
    import torch
    return {"f16": torch.float16, "bf16": torch.bfloat16, "f32": torch.float32}[dt]


def make_mpq_inputs(K, N, w_bit, group, dt, asym, M=1, act_order=False, seed=0, device="cpu"):
    import torch
    g = torch.Generator().manual_seed(1234 + seed)
    tdt = torch_dt(dt)
    nb = 32 // w_bit
    G = K // group
    qweight = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // nb, N), dtype=torch.int32, generator=g)
    scales = (torch.rand((G, N), generator=g) * 0.01 + 0.005).to(tdt)
    if asym:
        zeros = torch.randint(-2 ** 31, 2 ** 31 - 1, (G, N // nb), dtype=torch.int32, generator=g)
    else:
        zeros = (scales.float() * (2 ** (w_bit - 1)) + torch.randn((G, N), generator=g) * 1e-3).to(tdt)
    g_idx = torch.arange(K, dtype=torch.int32) // group
    if act_order:
        g_idx = g_idx[torch.randperm(K, generator=g)].contiguous()
    x = torch.randn((M, K), generator=g).to(tdt)
    out = dict(qweight=qweight, scales=scales, zeros=zeros, g_idx=g_idx, x=x)
    return {k: v.to(device) for k, v in out.items()}


def to_np_f32(t):
    """torch tensor (any float dtype) -> float32 numpy with exact values; ints pass through."""
    import torch
    t = t.detach().cpu()
    if t.dtype in (torch.float16, torch.bfloat16):
        return t.float().numpy()
    return t.numpy()


def rel_fro(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


# Stated tolerances (DESIGN.md "numerics").  north_star: "within 1e-3 relative for fp16 GEMM".
#   NORMWISE_TOL: ||y - y_ref||_F / ||y_ref||_F against the reference-faithful oracle (x.f32 @ unpack_qweight(q).f32).
#                 bf16 carries 8 significand bits, so its bound is two bf16 ulps, not 1e-3.
#   ELEM_RTOL   : element-wise bound against the exact-model oracle = output rounding (half ulp) + fp32 slack.
NORMWISE_TOL = {"f16": 1e-3, "bf16": 8e-3, "f32": 1e-4}
ELEM_RTOL = {"f16": 6e-4, "bf16": 4.5e-3, "f32": 2e-5}


def assert_close_to_oracles(y, y_ref, y_exact, dt, what=""):
    y = np.asarray(y, dtype=np.float64)
    nrm = rel_fro(y, y_ref)
    assert nrm <= NORMWISE_TOL[dt], f"{what}: normwise rel err {nrm:.3e} vs reference-faithful oracle > {NORMWISE_TOL[dt]}"
    rms = float(np.sqrt(np.mean(np.asarray(y_exact, dtype=np.float64) ** 2)))
    bound = ELEM_RTOL[dt] * np.abs(y_exact) + 1e-4 * rms
    worst = float(np.max(np.abs(y - y_exact) - bound))
    assert worst <= 0, f"{what}: element-wise error exceeds {ELEM_RTOL[dt]}*|y| + 1e-4*rms by {worst:.3e}"
    return nrm


# ---------------------------------------------------------------------------------------------------------------
# exl2 mixed-bit fixtures: restatement of the format helpers in the reference's tests
# (tests/layers/util.py:31-93 get_packed_info / get_q_groups), strategy of test_nbit_linear_mixbits.py:26-29
# ---------------------------------------------------------------------------------------------------------------
def exl2_packed_info(channels, n_bits, bits_prop, group_size):
    groups = rows = 0
    used = []
    sizes = list(group_size.values())
    for i in range(len(bits_prop)):
        gs = sizes[i]
        ch = max(1, int(channels * bits_prop[i]) // gs) * gs if i < len(bits_prop) - 1 else channels - sum(used)
        used.append(ch)
        groups += ch // gs
        rows += ch // 32 * n_bits[i]
    return groups, rows


def exl2_q_groups(groups, n_bits, group_size, channels, bits_prop):
    import math
    ends = []
    sizes = list(group_size.values())
    for i in range(len(bits_prop)):
        if i < len(bits_prop) - 1:
            e = max(1, int(channels * bits_prop[i]) // sizes[i]) * sizes[i] + (ends[-1] if ends else 0)
        else:
            e = channels
        ends.append(e)
    q = []
    for i, bits in enumerate(n_bits):
        rows = ends[i] - (ends[i - 1] if i else 0)
        q += [bits, 0] * (rows // group_size[str(bits)])
    out_row, rem = 0, channels
    for g in range(groups):
        bits = q[2 * g]
        gs = group_size[str(bits)]
        q[2 * g + 1] = out_row
        out_row += math.ceil(min(gs, rem) / (32 / bits))
    return q
