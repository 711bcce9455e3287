#define B200_PIPE_BITS 2
#include "mpq_pipe_family.inl"
