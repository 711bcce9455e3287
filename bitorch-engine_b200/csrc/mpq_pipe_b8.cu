#define B200_PIPE_BITS 8
#include "mpq_pipe_family.inl"
