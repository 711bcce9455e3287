#define B200_GEMV_BITS 4
#include "mpq_gemv_family.inl"
