#define B200_STREAM_BITS 4
#include "mpq_stream_family.inl"
