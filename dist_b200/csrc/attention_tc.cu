// placeholder until the tcgen05 attention kernel lands: bf16 attention runs on the SIMT kernel
#include "common.cuh"
namespace distb200 {
int attention_simt_launch(const void* qkv, void* out, int frames, int tokens, int heads, int dtype, cudaStream_t stream);
int attention_tc_launch(const void* qkv, void* out, int frames, int tokens, int heads, cudaStream_t stream) {
    return attention_simt_launch(qkv, out, frames, tokens, heads, DISTB200_BF16, stream);
}
}  // namespace distb200
