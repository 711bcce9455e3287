#define B200_STREAM_BITS 2
#include "mpq_stream_family.inl"
