"""Host logic of the decode chain (csrc/mpq_chain.cu): hazard analysis from the pointer ranges, strip deal, shadow ("LL")
assignment -- through b200bit_mpq_chain_plan_host, which needs no device.  Pointers are fake addresses: nothing is
dereferenced.  The dataflow is the one bench.py chains: a Llama decoder block q, k, v <- hidden; o <- v; gate, up <- o;
down <- up; next block <- down (reference call sequence: one q_linear_cuda.mpq_forward per layer, mpq_layer.py:65)."""
import ctypes

import pytest

from bitorch_engine_b200 import _cabi

H, I = 4096, 11008
BIT_RELEASE, BIT_HINT, BIT_SAME_X = 1 << 20, 1 << 21, 1 << 22


class _Arena:
    """fake, non-overlapping, 256-byte aligned device addresses"""
    def __init__(self):
        self.at = 0x10000000

    def take(self, nbytes):
        a = self.at
        self.at += (nbytes + 255) // 256 * 256
        return a


def _node(x, y, K, N, arena, G=None):
    G = G or K // 128
    n = _cabi.ChainNode()
    n.x, n.y, n.K, n.N, n.G = x, y, K, N, G
    n.qweight = arena.take(K // 8 * N * 4)
    n.scales = arena.take(G * N * 2)
    n.zeros = arena.take(G * N * 2)
    return n


def _plan(nodes, asym=0, dtype=None):
    lib = _cabi.lib()
    n = len(nodes)
    arr = (_cabi.ChainNode * n)(*nodes)
    table = (_cabi.ChainPlanNode * n)()
    info = (ctypes.c_int * 16)()
    rc = lib.b200bit_mpq_chain_plan_host(arr, n, 4, asym, _cabi.F16 if dtype is None else dtype, table, info)
    return rc, table, info


def _block(arena, hid):
    """seven nodes of one decoder block reading `hid`; returns (nodes, address of the block's output)"""
    q, k, v, o = (arena.take(H * 2) for _ in range(4))
    g, u = arena.take(I * 2), arena.take(I * 2)
    out = arena.take(H * 2)
    nodes = [_node(hid, q, H, H, arena), _node(hid, k, H, H, arena), _node(hid, v, H, H, arena), _node(v, o, H, H, arena),
             _node(o, g, H, I, arena), _node(o, u, H, I, arena), _node(u, out, I, H, arena)]
    return nodes, out


def test_decoder_block_dataflow():
    arena = _Arena()
    x0 = arena.take(H * 2)
    b0, out0 = _block(arena, x0)
    b1, _ = _block(arena, out0)
    rc, t, info = _plan(b0 + b1)
    assert rc == 0, _cabi.lib().b200bit_last_error()
    assert info[1] == 14 and info[2] == 148          # nodes, grid = one CTA per SM (148 assumed without a device)
    # first block: q, k, v read a buffer no node wrote: no wait, plain x; k and v re-read q's x
    for i in (0, 1, 2):
        assert t[i].wx_node == -1 and not t[i].xll
    assert not (t[0].off_sig & BIT_SAME_X) and (t[1].off_sig & BIT_SAME_X) and (t[2].off_sig & BIT_SAME_X)
    # o reads exactly v's output: through v's shadow, with v's counter as the hint
    assert t[3].wx_node == 2 and t[3].xll and t[3].xll == t[2].yll and (t[2].off_sig & BIT_HINT)
    # gate and up read o's output: both through ONE shadow; up is gate's sibling
    assert t[4].wx_node == 3 and t[5].wx_node == 3 and t[4].xll == t[3].yll == t[5].xll
    assert t[5].off_sig & BIT_SAME_X
    # down reads up's output (K = 11008)
    assert t[6].wx_node == 5 and t[6].xll == t[5].yll and t[6].R == I // 8 and t[6].tiles == 6
    # q and k feed nobody, gate feeds nobody: no shadow, no counter traffic
    for i in (0, 1, 4):
        assert not t[i].yll and not (t[i].off_sig & (BIT_RELEASE | BIT_HINT))
    # the next block reads the first block's output
    for i in (7, 8, 9):
        assert t[i].wx_node == 6 and t[i].xll == t[6].yll
    # no buffer is reused: nothing waits before writing y, nothing needs a release counter
    assert all(t[i].wy_node == -1 and not (t[i].off_sig & BIT_RELEASE) for i in range(14))
    # all shadows are distinct
    ll = [t[i].yll for i in range(14) if t[i].yll]
    assert len(ll) == len(set(ll)) == 7      # v, o, up, down of block 0; v, o, up of block 1 (its down feeds nobody)


def test_strip_deal():
    arena = _Arena()
    x0 = arena.take(H * 2)
    nodes, _ = _block(arena, x0)
    rc, t, info = _plan(nodes)
    assert rc == 0
    # 4096 columns: 146.3 strips of 28 -> 148 strips (136 x 28 + 12 x 24 columns), one per CTA; two 256-row tiles each
    assert (t[0].strips, t[0].n28, t[0].tiles) == (148, 136, 2)
    assert 136 * 28 + 12 * 24 == H
    # 11008 columns: 394 strips of 28 (rounding up to 444 would cost more than 3 % extra traffic)
    assert (t[4].strips, t[4].n28) == (394, 394)
    # the cyclic deal continues across nodes: node i starts where node i - 1 stopped (mod grid)
    off = 0
    for i in range(7):
        assert t[i].off_sig & 0xfffff == off
        off = (off + t[i].strips) % 148


def test_buffer_reuse_and_views_fall_back_to_counters():
    arena = _Arena()
    a, b = arena.take(H * 2), arena.take(H * 2)
    # ping-pong: a -> b -> a -> b.  Node 2 overwrites node 1's INPUT (write-after-read), node 3 overwrites node 1's output
    # after node 2 read it (write-after-read) -- both must wait on a counter before writing y
    nodes = [_node(a, b, H, H, arena), _node(b, a, H, H, arena), _node(a, b, H, H, arena)]
    rc, t, _ = _plan(nodes)
    assert rc == 0
    assert t[1].wx_node == 0 and t[1].xll                      # exact producer/consumer pair: shadow
    assert t[1].wy_node == -1 or t[1].wy_node <= t[1].wx_node  # writing a: node 0 read it, but the x wait already covers node 0
    assert t[2].wx_node == 1 and t[2].xll
    assert t[2].wy_node == -1 or t[2].wy_node <= 1             # writing b: node 1 read it; covered by the wait on node 1
    # a consumer that reads a SLICE of a producer's output is not an exact pair: ordered counter, release bit on the producer
    big = arena.take(2 * H * 2)
    nodes = [_node(a, big, H, 2 * H, arena), _node(big + H * 2, b, H, H, arena)]
    rc, t, _ = _plan(nodes)
    assert rc == 0
    assert t[1].wx_node == 0 and not t[1].xll and (t[0].off_sig & BIT_RELEASE) and not t[0].yll


def test_rejects_what_the_kernel_cannot_run():
    arena = _Arena()
    a, b = arena.take(H * 2), arena.take(H * 2)
    lib = _cabi.lib()
    rc, _, _ = _plan([_node(a, a, H, H, arena)])                       # writes its own input
    assert rc != 0 and b"own input" in lib.b200bit_last_error()
    rc, _, _ = _plan([_node(a, b, H, H, arena), _node(b, a, H, H, arena, G=H // 64)])      # two group sizes in one chain
    assert rc != 0 and b"group size" in lib.b200bit_last_error()
    rc, _, _ = _plan([_node(a, b, 4096 + 64, H, arena, G=1)])          # K not a multiple of 128
    assert rc != 0
    n = _node(a, b, H, H, arena)
    arr = (_cabi.ChainNode * 1)(n)
    table = (_cabi.ChainPlanNode * 1)()
    info = (ctypes.c_int * 16)()
    assert lib.b200bit_mpq_chain_plan_host(arr, 1, 2, 0, _cabi.F16, table, info) != 0     # the chain is the 4-bit kernel


def test_batched_kernel_support_predicate():
    """b200bit_mpq_forward_tc_supported is pure host arithmetic (tile height, split-K factor, group-table rows of a k-slice);
    the C dispatcher and the python shim both ask it.  148 SMs are assumed without a device."""
    lib = _cabi.lib()

    def ws(M, N):
        return _cabi.WS_TICKET_BYTES + 8 * min(M, 256) * N * 4 if M <= 256 else 0

    def ok(M, K, N, G, bits, asym=0, dtype=None, wsb=None):
        return lib.b200bit_mpq_forward_tc_supported(M, K, N, G, bits, asym, _cabi.F16 if dtype is None else dtype,
                                                    ws(M, N) if wsb is None else wsb)
    assert ok(32, 4096, 4096, 32, 4) and ok(2048, 4096, 11008, 32, 4) and ok(512, 14336, 4096, 112, 4)
    assert ok(32, 4096, 11008, 128, 2) and ok(33, 4096, 4096, 32, 4, asym=1)
    # 2-bit g32, K = 11008: 344 groups x 128 columns x (scale, zero) do not fit as a whole (176 KB) -- with split-K (workspace,
    # M <= 256) a slice touches 87 of them; without a workspace, or above 256 rows, the shape is declined
    assert ok(32, 11008, 4096, 344, 2) and ok(256, 11008, 4096, 344, 2)
    assert not ok(32, 11008, 4096, 344, 2, wsb=0) and not ok(512, 11008, 4096, 344, 2)
    # shape rules
    assert not ok(32, 4096, 4096, 32, 8) and not ok(32, 4096, 4096, 32, 4, dtype=_cabi.BF16)
    assert not ok(32, 4096 + 32, 4096, 1, 4) and not ok(32, 4096, 4100, 32, 4) and not ok(32, 4096, 4096, 4096 // 96, 4)
