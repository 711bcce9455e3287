#define B200_GEMV_BITS 8
#include "mpq_gemv_family.inl"
