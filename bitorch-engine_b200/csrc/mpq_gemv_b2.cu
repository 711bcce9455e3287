#define B200_GEMV_BITS 2
#include "mpq_gemv_family.inl"
