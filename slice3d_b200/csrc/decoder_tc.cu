// tcgen05 decoder -- placeholder until the fused kernel lands.
#include "common.cuh"

namespace s3d {

int dectc_pack(s3d_model*, cudaStream_t) { return S3D_OK; }
size_t decoder_tc_workspace_bytes(int64_t) { return 256; }
int decoder_tc(const s3d_model*, const float*, int, const QueryCtx&, int64_t, float, float*, int, void*, size_t,
               cudaStream_t) {
  set_error("decoder: tensor-core precision modes are not built in this revision");
  return S3D_ERR_UNSUPPORTED;
}

}  // namespace s3d
