#define B200_STREAM_BITS 8
#include "mpq_stream_family.inl"
