#define B200_MMA_BITS 8
#include "mpq_mma_family.inl"
