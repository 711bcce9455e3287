#define B200_PIPE_BITS 4
#include "mpq_pipe_family.inl"
