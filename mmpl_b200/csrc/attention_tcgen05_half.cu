// The half-tile S pipeline variant of flash_attn_kernel (MMPL_ATTN_SPLIT_S = 1) as namespace mmpl::half: opt-in through
// the dispatcher (attention_dispatch.cu: MMPL_ATTN_HALF=1 / MMPL_ATTN_HALF_TILES=n), see the note at
// MMPL_ATTN_SPLIT_S in attention_tcgen05.cu.
#define MMPL_ATTN_NS half
#define MMPL_ATTN_SPLIT_S 1
#include "attention_tcgen05.cu"
