#define B200_MMA_BITS 2
#include "mpq_mma_family.inl"
