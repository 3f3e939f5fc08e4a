// The half-tile S pipeline variant of flash_attn_kernel (MMPL_ATTN_SPLIT_S = 1) as namespace mmpl::half: used by the
// dispatcher (attention_dispatch.cu) for short KV ranges (cross-attention over the 512 text keys), see the note at
// MMPL_ATTN_SPLIT_S in attention_tcgen05.cu.
#define MMPL_ATTN_NS half
#define MMPL_ATTN_SPLIT_S 1
#include "attention_tcgen05.cu"
