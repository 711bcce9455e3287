"""TEST INFRASTRUCTURE ONLY.

CPU restatement ("oracle") of the reference's low-bit Linear hot path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this package, and only
as the checker / the CPU baseline -- never as the thing measured or shipped.  The product package
(`bitorch-engine_b200/`) never imports it and fails loudly when its CUDA library is missing.

Pinning (see DESIGN.md "Oracle"):
  * n-bit unpack / pack / dequant / forward / grad_input / optimizer update: pinned against the reference's own
    Python (`/root/reference/bitorch_engine/...`) executed in the build container by `oracle/gen_golden.py`; the
    resulting vectors are committed under `tests/golden/`.
  * binary linear: pinned against `oracle/_ref/binary_linear_ref.so`, compiled unmodified from
    `/root/reference/bitorch_engine/layers/qlinear/binary/cpp/binary_linear.cpp` by `oracle/build_ref.py`,
    and against the golden vectors generated with it.
"""
