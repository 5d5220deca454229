// tcgen05 flash attention (placeholder until the TMEM softmax pipeline lands; see attention.cu for the SIMT path)
#include "common.cuh"
#include "ttb_internal.h"
namespace ttb {
bool flash_attention_supported(const TtbAttnArgs&) { return false; }
int flash_attention_launch(const TtbAttnArgs&, cudaStream_t) { set_error("flash attention not built"); return -1; }
}
