#define B200_MMA_BITS 4
#include "mpq_mma_family.inl"
