from bitorch_engine_b200.extensions.q_linear_cuda import *  # noqa: F401,F403
from bitorch_engine_b200.extensions.q_linear_cuda import mpq_forward, mpq_grad_input, mbwq_trans_qweight, mbwq_q42fp_weight, mbwq_q4_forward, mbwq_exl2fp_weight, mbwq_exl2_forward  # noqa: F401
