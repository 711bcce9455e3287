#define B200_GEMV_BITS 1
#include "mpq_gemv_family.inl"
