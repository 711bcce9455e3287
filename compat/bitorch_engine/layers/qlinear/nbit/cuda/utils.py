from bitorch_engine_b200.layers.qlinear.nbit.cuda.utils import unpack_qweight, pack_fp_weight, make_group_map  # noqa: F401
