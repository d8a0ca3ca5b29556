// C ABI of slice3d_b200 (include/slice3d_b200.h): model creation (weight folding / packing)
// and the thin entry points over encoder.cu / decoder_simt.cu / decoder_tc.cu.
#include <cmath>
#include <cstring>
#include <map>

#include "common.cuh"

namespace s3d {

static thread_local std::string g_err;
std::atomic<long long> g_launches{0};
void set_error(const std::string& msg) { g_err = msg; }

namespace {

constexpr float kBnEps = 1e-5f;

struct Packer {
  s3d_model* m;
  std::map<std::string, const s3d_tensor*> by_name;
  cudaStream_t st;
  std::string err;

  bool fetch(const std::string& name, int64_t numel, std::vector<float>& out) {
    auto it = by_name.find(name);
    if (it == by_name.end()) {
      err = "missing tensor: " + name;
      return false;
    }
    if (it->second->numel != numel) {
      err = "tensor " + name + ": numel " + std::to_string(it->second->numel) + ", expected " + std::to_string(numel);
      return false;
    }
    out.resize((size_t)numel);
    cudaError_t e = cudaMemcpyAsync(out.data(), it->second->data_dev, (size_t)numel * sizeof(float),
                                    cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      err = "copy of " + name + " failed: " + cudaGetErrorString(e);
      return false;
    }
    return true;
  }

  float* upload(const std::vector<float>& h) {
    float* d = nullptr;
    if (cudaMalloc(&d, h.size() * sizeof(float)) != cudaSuccess) {
      err = "cudaMalloc failed";
      return nullptr;
    }
    m->allocs.push_back(d);
    if (cudaMemcpyAsync(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {
      err = "upload failed";
      return nullptr;
    }
    return d;
  }

  // eval-mode BatchNorm as y = x*scale + shift; `bias` (may be empty) is the preceding
  // convolution's bias folded into shift.
  bool bn_fold(const std::string& p, int C, const std::vector<float>& bias, std::vector<float>& scale,
               std::vector<float>& shift) {
    std::vector<float> g, b, mu, var;
    if (!fetch(p + ".weight", C, g) || !fetch(p + ".bias", C, b) || !fetch(p + ".running_mean", C, mu) ||
        !fetch(p + ".running_var", C, var))
      return false;
    scale.resize(C);
    shift.resize(C);
    for (int c = 0; c < C; ++c) {
      double s = (double)g[c] / std::sqrt((double)var[c] + (double)kBnEps);
      double cb = bias.empty() ? 0.0 : (double)bias[c];
      scale[c] = (float)s;
      shift[c] = (float)((cb - (double)mu[c]) * s + (double)b[c]);
    }
    return true;
  }

  // Conv2d weight (Cout, Cin, ks, ks) -> [(tap*CinP + ci)][Cout], CinP = Cin rounded up to 4.
  bool pack_conv(const std::string& wname, int Cout, int Cin, int ks, ConvW& cw) {
    std::vector<float> w;
    if (!fetch(wname, (int64_t)Cout * Cin * ks * ks, w)) return false;
    const int CinP = (Cin + 3) / 4 * 4, taps = ks * ks;
    cw.cin = CinP;
    cw.ncols = Cout;
    cw.ks = ks;
    cw.k = taps * CinP;
    cw.kpad = (cw.k + 15) / 16 * 16;
    std::vector<float> t((size_t)cw.kpad * Cout, 0.f);
    for (int co = 0; co < Cout; ++co)
      for (int ci = 0; ci < Cin; ++ci)
        for (int tp = 0; tp < taps; ++tp)
          t[((size_t)tp * CinP + ci) * Cout + co] = w[((size_t)co * Cin + ci) * taps + tp];
    cw.w = upload(t);
    return cw.w != nullptr;
  }

  // Linear weight (out, in) restricted to input columns [c0, c0+cin) -> [cin][out]
  bool pack_linear(const std::string& wname, int out, int in, int c0, int cin, ConvW& cw) {
    std::vector<float> w;
    if (!fetch(wname, (int64_t)out * in, w)) return false;
    cw.cin = cin;
    cw.ncols = out;
    cw.ks = 1;
    cw.k = cin;
    cw.kpad = (cin + 15) / 16 * 16;
    std::vector<float> t((size_t)cw.kpad * out, 0.f);
    for (int o = 0; o < out; ++o)
      for (int i = 0; i < cin; ++i) t[(size_t)i * out + o] = w[(size_t)o * in + c0 + i];
    cw.w = upload(t);
    return cw.w != nullptr;
  }

  bool vec(const std::string& name, int n, float*& dst) {
    std::vector<float> v;
    if (!fetch(name, n, v)) return false;
    dst = upload(v);
    return dst != nullptr;
  }

  // ---- VGG19 perceptual loss (optional; vgg_perceptual_loss.py:6-40: torchvision feature indices per slice), plus the
  //      weights of the data-gradient convolutions of its backward pass: dX = conv3x3(dY, W') with
  //      W'[ci][co][ky][kx] = W[co][ci][2-ky][2-kx] (the network is frozen: no weight gradients)
  bool build_pvgg() {
    struct PC {
      int slice, idx, cin, cout;
    };
    const PC pc[14] = {{1, 0, 3, 64},     {1, 2, 64, 64},    {2, 5, 64, 128},   {2, 7, 128, 128},  {3, 10, 128, 256},
                       {3, 12, 256, 256}, {4, 14, 256, 256}, {4, 16, 256, 256}, {4, 19, 256, 512}, {4, 21, 512, 512},
                       {5, 23, 512, 512}, {5, 25, 512, 512}, {5, 28, 512, 512}, {5, 30, 512, 512}};
    for (int i = 0; i < 14; ++i) {
      std::string p = "vggptlossfunc.vgg.slice" + std::to_string(pc[i].slice) + "." + std::to_string(pc[i].idx);
      if (!pack_conv(p + ".weight", pc[i].cout, pc[i].cin, 3, m->pvgg[i])) return false;
      if (!vec(p + ".bias", pc[i].cout, m->pvgg[i].shift)) return false;
      std::vector<float> w;
      const int Cout = pc[i].cout, Cin = pc[i].cin, CinP = (Cin + 3) / 4 * 4;
      if (!fetch(p + ".weight", (int64_t)Cout * Cin * 9, w)) return false;
      ConvW& d = m->pvgg_d[i];
      d.cin = Cout; d.ncols = CinP; d.ks = 3; d.k = 9 * Cout; d.kpad = (d.k + 15) / 16 * 16;
      std::vector<float> t((size_t)d.kpad * CinP, 0.f);
      for (int co = 0; co < Cout; ++co)
        for (int ci = 0; ci < Cin; ++ci)
          for (int tp = 0; tp < 9; ++tp) t[((size_t)tp * Cout + co) * CinP + ci] = w[((size_t)co * Cin + ci) * 9 + (8 - tp)];
      if (!(d.w = upload(t))) return false;
    }
    if (!vec("vggptlossfunc.mean", 3, m->pvgg_mean) || !vec("vggptlossfunc.std", 3, m->pvgg_std)) return false;
    m->has_pvgg = 1;
    return true;
  }

  bool build() {
    if (!by_name.count("slices_generator.down1.0.weight") && !by_name.count("img_encoder.conv1_2.0.weight") &&
        by_name.count("vggptlossfunc.vgg.slice1.0.weight")) {
      m->kind = 2;  // the frozen perceptual network alone (training: its weights never change, the model's do every step)
      return build_pvgg();
    }
    // Slices3DGTModel checkpoints (reg_slices/src/model_gt.py) are recognised by their encoder keys: the same VGG16-BN
    // trunk under img_encoder.conv1_2 .. conv5_3 (vgg16bn_feats.py:27-32: same cut points as down1 .. down5)
    const bool gt = by_name.count("img_encoder.conv1_2.0.weight") != 0;
    m->kind = gt ? 1 : 0;
    const std::string g = gt ? "img_encoder." : "slices_generator.";
    auto blk_name = [&](const char* down) -> std::string {
      if (!gt) return down;
      static const std::map<std::string, std::string> mp = {{"down1", "conv1_2"}, {"down2", "conv2_2"}, {"down3", "conv3_3"},
                                                            {"down4", "conv4_3"}, {"down5", "conv5_3"}};
      return mp.at(down);
    };
    // ---- VGG16-BN trunk (unet_custom.py:15-19; torchvision feature indices)
    struct VC {
      const char* blk;
      int idx, cin, cout, bn;  // bn = index of the BatchNorm that follows inside the block, -1 for the tap conv
    };
    const VC vc[13] = {{"down1", 0, 3, 64, 1},     {"down1", 3, 64, 64, -1},   {"down2", 7, 64, 128, 8},
                       {"down2", 10, 128, 128, -1}, {"down3", 14, 128, 256, 15}, {"down3", 17, 256, 256, 18},
                       {"down3", 20, 256, 256, -1}, {"down4", 24, 256, 512, 25}, {"down4", 27, 512, 512, 28},
                       {"down4", 30, 512, 512, -1}, {"down5", 34, 512, 512, 35}, {"down5", 37, 512, 512, 38},
                       {"down5", 40, 512, 512, -1}};
    for (int i = 0; i < 13; ++i) {
      std::string p = g + blk_name(vc[i].blk) + "." + std::to_string(vc[i].idx);
      if (!pack_conv(p + ".weight", vc[i].cout, vc[i].cin, 3, m->vgg[i])) return false;
      std::vector<float> bias;
      if (!fetch(p + ".bias", vc[i].cout, bias)) return false;
      if (vc[i].bn >= 0) {
        std::vector<float> sc, sh;
        if (!bn_fold(g + blk_name(vc[i].blk) + "." + std::to_string(vc[i].bn), vc[i].cout, bias, sc, sh)) return false;
        if (!(m->vgg[i].scale = upload(sc)) || !(m->vgg[i].shift = upload(sh))) return false;
      } else {
        if (!(m->vgg[i].shift = upload(bias))) return false;
      }
    }
    const char* leadb[4] = {"down2", "down3", "down4", "down5"};
    const char* leadi[4] = {".4", ".11", ".21", ".31"};
    const int leadc[4] = {64, 128, 256, 512};
    for (int i = 0; i < 4; ++i) {
      std::vector<float> sc, sh;
      if (!bn_fold(g + blk_name(leadb[i]) + leadi[i], leadc[i], {}, sc, sh)) return false;
      if (!(m->bn_scale[i] = upload(sc)) || !(m->bn_shift[i] = upload(sh))) return false;
    }
    DecF32& d = m->dec32;
    if (gt) {
      // ---- fc_local (model_gt.py:33-38): the first Linear(1472, 128) hoisted onto the taps.  Its input channels follow the
      //      concatenation order conv1_2 (64) .. conv5_3 (512); tap i has resolution S / 2^i = plane scale 4 - i of the blob.
      const int tapc[5] = {64, 128, 256, 512, 512};
      int c0 = 0;
      for (int i = 0; i < 5; ++i) {
        if (!pack_linear("fc_local.0.weight", 128, 1472, c0, tapc[i], m->fcs[4 - i])) return false;
        c0 += tapc[i];
      }
      if (!vec("fc_local.0.bias", 128, d.fcs_b)) return false;
      if (!pack_linear("fc_local.2.weight", 128, 128, 0, 128, m->fcl2) || !vec("fc_local.2.bias", 128, m->fcl2.shift)) return false;
      // ---- pts_feat_extractor (model_gt.py:24-31): raw PyTorch (out, in) matrices, evaluated per query by one warp
      if (!vec("pts_feat_extractor.0.weight", 32 * 3, m->pts_w[0]) || !vec("pts_feat_extractor.0.bias", 32, m->pts_b[0]) ||
          !vec("pts_feat_extractor.2.weight", 64 * 32, m->pts_w[1]) || !vec("pts_feat_extractor.2.bias", 64, m->pts_b[1]) ||
          !vec("pts_feat_extractor.4.weight", 128 * 64, m->pts_w[2]) || !vec("pts_feat_extractor.4.bias", 128, m->pts_b[2]))
        return false;
    } else {
    // ---- trans_c: 1x1 conv over cat[x5 (512), slice embedding (128)] (unet_custom.py:52-57)
    {
      const int K = m->K;
      std::vector<float> w, b, emb;
      if (!fetch(g + "trans_c.weight", 512 * 640, w) || !fetch(g + "trans_c.bias", 512, b) ||
          !fetch(g + "emds.weight", (int64_t)K * 128, emb))
        return false;
      ConvW& cw = m->trans_c;
      cw.cin = 512; cw.ncols = 512; cw.ks = 1; cw.k = 512; cw.kpad = 512;
      std::vector<float> t((size_t)512 * 512);
      for (int co = 0; co < 512; ++co)
        for (int ci = 0; ci < 512; ++ci) t[(size_t)ci * 512 + co] = w[(size_t)co * 640 + ci];
      if (!(cw.w = upload(t))) return false;
      std::vector<float> e((size_t)K * 512);
      for (int k = 0; k < K; ++k)
        for (int co = 0; co < 512; ++co) {
          float s = 0.f;  // fp32 like the reference's convolution
          for (int j = 0; j < 128; ++j) s = fmaf(w[(size_t)co * 640 + 512 + j], emb[(size_t)k * 128 + j], s);
          e[(size_t)k * 512 + co] = s + b[co];
        }
      if (!(m->trans_c_e = upload(e))) return false;
    }
    // ---- Up stages
    for (int n = 1; n <= 4; ++n) {
      const int Cin = kPlaneC[n - 1], C = kPlaneC[n];
      std::string up = g + "up" + std::to_string(n);
      {  // ConvTranspose2d weight (Cin, C, 2, 2) -> [ci][(dy*2+dx)*C + co]
        std::vector<float> w, b;
        if (!fetch(up + ".up.weight", (int64_t)Cin * C * 4, w) || !fetch(up + ".up.bias", C, b)) return false;
        ConvW& cw = m->up_t[n - 1];
        cw.cin = Cin; cw.ncols = 4 * C; cw.ks = 1; cw.k = Cin; cw.kpad = Cin;
        std::vector<float> t((size_t)Cin * 4 * C);
        for (int ci = 0; ci < Cin; ++ci)
          for (int co = 0; co < C; ++co)
            for (int q = 0; q < 4; ++q) t[(size_t)ci * 4 * C + (size_t)q * C + co] = w[((size_t)ci * C + co) * 4 + q];
        if (!(cw.w = upload(t)) || !(cw.shift = upload(b))) return false;
      }
      const std::string dc = up + ".conv.double_conv";
      if (!pack_conv(dc + ".0.weight", C, 2 * C, 3, m->dc1[n - 1])) return false;
      if (!pack_conv(dc + ".3.weight", C, C, 3, m->dc2[n - 1])) return false;
      std::vector<float> sc, sh;
      if (!bn_fold(dc + ".1", C, {}, sc, sh)) return false;
      if (!(m->dc1[n - 1].scale = upload(sc)) || !(m->dc1[n - 1].shift = upload(sh))) return false;
      if (!bn_fold(dc + ".4", C, {}, sc, sh)) return false;
      if (!(m->dc2[n - 1].scale = upload(sc)) || !(m->dc2[n - 1].shift = upload(sh))) return false;
      // 1x1 skip adapter trans_up{n}: 2C -> C
      std::string tu = g + "trans_up" + std::to_string(n);
      if (!pack_conv(tu + ".weight", C, 2 * C, 1, m->trans_up[n - 1])) return false;
      if (!vec(tu + ".bias", C, m->trans_up[n - 1].shift)) return false;
    }
    if (!vec(g + "outc.conv.weight", 96, m->outc_w) || !vec(g + "outc.conv.bias", 3, m->outc_b)) return false;
    // ---- fc_s hoisted per scale (models.py:80; input channel order = scale order 512,256,128,64,32)
    {
      int c0 = 0;
      for (int s = 0; s < 5; ++s) {
        if (!pack_linear("fc_s.weight", 128, 992, c0, kPlaneC[s], m->fcs[s])) return false;
        c0 += kPlaneC[s];
      }
    }
    if (by_name.count("vggptlossfunc.vgg.slice1.0.weight") && !build_pvgg()) return false;
    // ---- decoder
    {
      std::vector<float> w;
      if (!fetch("fc_p.weight", 128 * 3, w)) return false;
      std::vector<float> t(3 * 128);
      for (int c = 0; c < 128; ++c)
        for (int j = 0; j < 3; ++j) t[j * 128 + c] = w[c * 3 + j];
      if (!(d.fcp_wt = upload(t))) return false;
    }
    if (!vec("fc_p.bias", 128, d.fcp_b) || !vec("fc_s.bias", 128, d.fcs_b)) return false;
    }  // !gt
    if (!vec("fc_out.0.weight", 128, d.fco_w) || !vec("fc_out.0.bias", 1, d.fco_b)) return false;
    for (int l = 0; l < 3; ++l) {
      std::string p = "att_decoder.layers." + std::to_string(l);
      DecLayerF32& L = d.L[l];
      if (!pack_linear(p + ".self_attn.in_proj_weight", 384, 128, 0, 128, L.in_proj) ||
          !vec(p + ".self_attn.in_proj_bias", 384, L.in_proj.shift) ||
          !pack_linear(p + ".self_attn.out_proj.weight", 128, 128, 0, 128, L.out_proj) ||
          !vec(p + ".self_attn.out_proj.bias", 128, L.out_proj.shift) ||
          !pack_linear(p + ".linear1.weight", 2048, 128, 0, 128, L.lin1) || !vec(p + ".linear1.bias", 2048, L.lin1.shift) ||
          !pack_linear(p + ".linear2.weight", 128, 2048, 0, 2048, L.lin2) || !vec(p + ".linear2.bias", 128, L.lin2.shift) ||
          !vec(p + ".norm1.weight", 128, L.n1_w) || !vec(p + ".norm1.bias", 128, L.n1_b) ||
          !vec(p + ".norm2.weight", 128, L.n2_w) || !vec(p + ".norm2.bias", 128, L.n2_b))
        return false;
    }
    return true;
  }
};

__global__ void k_flip_yz(float* q, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  q[3 * i + 1] = -q[3 * i + 1];
  q[3 * i + 2] = -q[3 * i + 2];
}

// Argument / capability / workspace validation of a decoder call, separate from the launch so that entry points with
// side effects on the caller's buffers (the in-place y,z flip) can validate first.
int check_decoder(const s3d_model* m, const void* planes, int S, const float* T, int64_t n, const float* out, int precision,
                  const float* debug_tokens, const void* ws, size_t ws_bytes, bool gt_tokens = false) {
  if (!m || !planes || !out || !T || n < 0 || S < 32 || S % 16) {
    set_error("decoder: bad argument");
    return S3D_ERR_BAD_ARG;
  }
  if ((m->kind != 0 && !gt_tokens) || m->kind == 2) {
    set_error("decoder: this handle holds a Slices3DGTModel (use s3d_gt_decoder_fwd) or only the perceptual network");
    return S3D_ERR_BAD_ARG;
  }
  if (precision != S3D_PREC_FP32 && precision != S3D_PREC_BF16X3 && precision != S3D_PREC_BF16 &&
      precision != S3D_PREC_FP16X3 && precision != S3D_PREC_FP16F8) {
    set_error("decoder: unknown precision mode");
    return S3D_ERR_BAD_ARG;
  }
  if (precision != S3D_PREC_FP32) {
    if (debug_tokens) {
      set_error("decoder: token dump is only available with S3D_PREC_FP32");
      return S3D_ERR_UNSUPPORTED;
    }
    if (!decoder_tc_supported(m)) {
      set_error("decoder: this n_slices has no tensor-core instantiation (use S3D_PREC_FP32)");
      return S3D_ERR_UNSUPPORTED;
    }
  } else if (m->K > 12) {
    set_error("decoder(fp32): n_slices > 12 unsupported");
    return S3D_ERR_UNSUPPORTED;
  }
  const size_t need = gt_tokens ? gt_decoder_workspace_bytes(n, precision) : s3d_decoder_workspace_bytes(n, precision);
  if (n > 0 && (ws == nullptr || ws_bytes < need)) {
    set_error("decoder: workspace too small");
    return S3D_ERR_WORKSPACE;
  }
  return S3D_OK;
}

int run_decoder(const s3d_model* m, const void* planes, int S, const QueryCtx& q, int64_t n, float out_scale,
                float* out, int precision, float* debug_tokens, void* ws, size_t ws_bytes, cudaStream_t st) {
  S3D_TRY(check_decoder(m, planes, S, q.T, n, out, precision, debug_tokens, ws, ws_bytes));
  if (precision == S3D_PREC_FP32)
    return decoder_simt(m, static_cast<const float*>(planes), S, q, n, out_scale, out, debug_tokens, ws, ws_bytes, st);
  return decoder_tc(m, static_cast<const float*>(planes), S, q, n, out_scale, out, precision, ws, ws_bytes, st);
}

}  // namespace
}  // namespace s3d

using namespace s3d;

extern "C" {

int s3d_abi_version(void) { return S3D_ABI_VERSION; }
const char* s3d_last_error(void) { return g_err.c_str(); }
int64_t s3d_launch_count(void) { return g_launches.load(); }

int s3d_model_create(s3d_model** out, const s3d_tensor* tensors, int32_t n_tensors, int32_t n_slices, int32_t device,
                     void* stream) {
  if (!out || !tensors || n_tensors <= 0 || n_slices < 1 || n_slices > 12) {
    set_error("model_create: bad argument (n_slices must be 1..12)");
    return S3D_ERR_BAD_ARG;
  }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("model_create: no CUDA device; slice3d_b200 has no CPU path");
    return S3D_ERR_CUDA;
  }
  S3D_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  S3D_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("model_create: this library is built for sm_100a (B200) only");
    return S3D_ERR_UNSUPPORTED;
  }
  s3d_model* m = new s3d_model();
  m->device = device;
  m->K = n_slices;
  Packer pk{m, {}, static_cast<cudaStream_t>(stream), ""};
  for (int i = 0; i < n_tensors; ++i)
    if (tensors[i].name) pk.by_name[tensors[i].name] = &tensors[i];
  if (!pk.build()) {
    set_error("model_create: " + pk.err);
    bool missing = pk.err.rfind("missing", 0) == 0;
    s3d_model_destroy(m);
    return missing ? S3D_ERR_MISSING_TENSOR : S3D_ERR_BAD_ARG;
  }
  int r = enctc_pack(m, static_cast<cudaStream_t>(stream));
  if (r == S3D_OK && m->kind != 2) r = dectc_pack(m, static_cast<cudaStream_t>(stream));
  if (r != S3D_OK) {
    s3d_model_destroy(m);
    return r;
  }
  *out = m;
  return S3D_OK;
}

void s3d_model_destroy(s3d_model* m) {
  if (!m) return;
  for (void* p : m->allocs) cudaFree(p);
  delete m;
}

int s3d_model_n_slices(const s3d_model* m) { return m ? m->K : 0; }

size_t s3d_planes_bytes(int32_t B, int32_t K, int32_t S) {
  if (B <= 0 || K <= 0 || S <= 0) return 0;
  return (size_t)B * plane_offset_floats(K, S, 5) * sizeof(float);
}

size_t s3d_encoder_workspace_bytes(int32_t B, int32_t K, int32_t S) {
  if (B <= 0 || K <= 0 || S <= 0) return 0;
  return encoder_workspace_bytes(B, K, S);
}

int s3d_encoder_fwd(const s3d_model* m, const float* img_dev, int32_t B, int32_t S, void* planes_dev,
                    float* const* feats_nchw_dev, float* slices_rec_dev, void* workspace_dev, size_t workspace_bytes,
                    void* stream) {
  if (!m || !img_dev) {
    set_error("encoder: null model or image");
    return S3D_ERR_BAD_ARG;
  }
  return encoder_fwd(m, img_dev, B, S, planes_dev, feats_nchw_dev, slices_rec_dev, workspace_dev, workspace_bytes,
                     static_cast<cudaStream_t>(stream));
}

size_t s3d_decoder_workspace_bytes(int64_t n, int32_t precision) {
  if (precision == S3D_PREC_FP32) return decoder_simt_workspace_bytes(n);
  return decoder_tc_workspace_bytes(n);
}

int s3d_decoder_batch_fwd(const s3d_model* m, const void* planes_dev, int32_t S, float* qry_dev, int32_t B,
                          int64_t n_per_image, const float* T_dev, const float* rot_dev, int32_t flip_in_place,
                          float out_scale, float* out_dev, int32_t precision, void* workspace_dev, size_t workspace_bytes,
                          void* stream) {
  if (B < 1 || n_per_image < 0) {
    set_error("decoder: bad batch");
    return S3D_ERR_BAD_ARG;
  }
  const int64_t n = (int64_t)B * n_per_image;
  if (n == 0) return S3D_OK;  /* empty query set (a zero-element tensor has a null data pointer) */
  if (!qry_dev) {
    set_error("decoder: null query pointer");
    return S3D_ERR_BAD_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  QueryCtx q{};
  q.qry = qry_dev;
  q.T = T_dev;
  q.rot = rot_dev;
  if (B > 1) {
    q.per_img = n_per_image;
    q.plane_stride = plane_offset_floats(m ? m->K : 0, S, 5);
  }
  // validate everything before touching the caller's query tensor: a failing call must leave it unflipped
  S3D_TRY(check_decoder(m, planes_dev, S, T_dev, n, out_dev, precision, nullptr, workspace_dev, workspace_bytes));
  if (!rot_dev && flip_in_place && n > 0) {
    k_flip_yz<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(qry_dev, n);
    S3D_LAUNCH_CHECK();
    q.preflipped = 1;
  }
  return run_decoder(m, planes_dev, S, q, n, out_scale, out_dev, precision, nullptr, workspace_dev, workspace_bytes, st);
}

int s3d_decoder_fwd(const s3d_model* m, const void* planes_dev, int32_t S, float* qry_dev, int64_t n,
                    const float* T_dev, const float* rot_dev, int32_t flip_in_place, float out_scale, float* out_dev,
                    int32_t precision, void* workspace_dev, size_t workspace_bytes, void* stream) {
  return s3d_decoder_batch_fwd(m, planes_dev, S, qry_dev, 1, n, T_dev, rot_dev, flip_in_place, out_scale, out_dev,
                               precision, workspace_dev, workspace_bytes, stream);
}

int s3d_decoder_grid_fwd(const s3d_model* m, const void* planes_dev, int32_t S, const s3d_grid* grid, int64_t first,
                         int64_t count, const float* T_dev, float out_scale, float* out_dev, int32_t precision,
                         void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (!grid || !grid->px_dev || !grid->py_dev || !grid->pz_dev || grid->nx <= 0 || grid->ny <= 0 || grid->nz <= 0 ||
      first < 0 || count < 0 || first + count > (int64_t)grid->nx * grid->ny * grid->nz) {
    set_error("decoder_grid: bad grid descriptor or range");
    return S3D_ERR_BAD_ARG;
  }
  QueryCtx q{};
  q.nx = grid->nx; q.ny = grid->ny; q.nz = grid->nz;
  q.px = grid->px_dev; q.py = grid->py_dev; q.pz = grid->pz_dev;
  q.first = first;
  q.T = T_dev;
  const int64_t plane = (int64_t)grid->ny * grid->nz;
  if (precision != S3D_PREC_FP32 && first % plane == 0 && count % plane == 0 && count > 0) {
    q.blk = 1;  // whole x-planes: evaluate in the locality order of common.cuh (the output layout is unchanged)
    q.x0 = (int)(first / plane);
    q.nxs = (int)(count / plane);
  }
  return run_decoder(m, planes_dev, S, q, count, out_scale, out_dev, precision, nullptr, workspace_dev,
                     workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t s3d_vgg_loss_workspace_bytes(int32_t N, int32_t S) {
  if (N <= 0 || S <= 0) return 0;
  return vgg_loss_workspace_bytes(N, S);
}

int s3d_vgg_loss_fwd(const s3d_model* m, const float* a_dev, const float* b_dev, int32_t N, int32_t S, float* loss_dev,
                     void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (!m) {
    set_error("vgg_loss: null model");
    return S3D_ERR_BAD_ARG;
  }
  return vgg_loss_fwd(m, a_dev, b_dev, N, S, loss_dev, workspace_dev, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int s3d_mc_count(const double* vol_dev, int32_t nx, int32_t ny, int32_t nz, double isovalue, const int32_t* tri_count_dev,
                 int32_t* vcount_dev, int32_t* tcount_dev, uint8_t* owned_dev, void* stream) {
  return mc_count(vol_dev, nx, ny, nz, isovalue, tri_count_dev, vcount_dev, tcount_dev, owned_dev,
                  static_cast<cudaStream_t>(stream));
}

int s3d_mc_emit(const double* vol_dev, int32_t nx, int32_t ny, int32_t nz, double isovalue, const int8_t* table_dev,
                const int64_t* vbase_dev, const int64_t* tbase_dev, const int32_t* tcount_dev, const uint8_t* owned_dev,
                double* verts_dev, int64_t* tris_dev, void* stream) {
  return mc_emit(vol_dev, nx, ny, nz, isovalue, reinterpret_cast<const signed char*>(table_dev),
                 reinterpret_cast<const long long*>(vbase_dev), reinterpret_cast<const long long*>(tbase_dev), tcount_dev,
                 owned_dev, verts_dev, reinterpret_cast<long long*>(tris_dev), static_cast<cudaStream_t>(stream));
}

size_t s3d_train_decoder_saved_bytes(const s3d_train_cfg* cfg) { return train_decoder_saved_bytes(cfg); }
size_t s3d_train_decoder_bwd_workspace_bytes(const s3d_train_cfg* cfg) { return train_decoder_bwd_workspace_bytes(cfg); }
int s3d_train_decoder_fwd(const s3d_train_cfg* cfg, const float* const* feats_dev, const float* qry_dev, const float* T_dev,
                          const float* const* params_dev, float* sdf_dev, void* saved_dev, size_t saved_bytes, void* stream) {
  return train_decoder_fwd(cfg, feats_dev, qry_dev, T_dev, params_dev, sdf_dev, saved_dev, saved_bytes,
                           static_cast<cudaStream_t>(stream));
}
int s3d_train_decoder_bwd(const s3d_train_cfg* cfg, const float* qry_dev, const float* T_dev, const float* const* params_dev,
                          const float* dsdf_dev, void* saved_dev, size_t saved_bytes, float* const* dfeats_dev,
                          float* const* dparams_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
  return train_decoder_bwd(cfg, qry_dev, T_dev, params_dev, dsdf_dev, saved_dev, saved_bytes, dfeats_dev, dparams_dev,
                           workspace_dev, workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t s3d_mise_scratch_ints(int32_t resolution0, int32_t depth) { return mise_scratch_ints(resolution0, depth); }

int s3d_mise_subdivide(int32_t resolution0, int32_t depth, double threshold, const double* value_dev, const uint8_t* known_dev,
                       int8_t* cell_level_dev, uint8_t* exists_dev, int32_t* flags_dev, void* stream) {
  return mise_subdivide(resolution0, depth, threshold, value_dev, known_dev, reinterpret_cast<signed char*>(cell_level_dev),
                        exists_dev, flags_dev, static_cast<cudaStream_t>(stream));
}

size_t s3d_vgg_loss_train_bytes(int32_t N, int32_t S) {
  if (N <= 0 || S <= 0) return 0;
  return vgg_loss_train_bytes(N, S);
}
int s3d_vgg_loss_train_fwd(const s3d_model* m, const float* a_dev, const float* b_dev, int32_t N, int32_t S, float* loss_dev,
                           void* saved_dev, size_t saved_bytes, void* stream) {
  if (!m) {
    set_error("vgg_loss_train: null model");
    return S3D_ERR_BAD_ARG;
  }
  return vgg_loss_train_fwd(m, a_dev, b_dev, N, S, loss_dev, saved_dev, saved_bytes, static_cast<cudaStream_t>(stream));
}
int s3d_vgg_loss_train_bwd(const s3d_model* m, int32_t N, int32_t S, const float* gout_dev, void* saved_dev, size_t saved_bytes,
                           float* grad_a_dev, void* stream) {
  if (!m) {
    set_error("vgg_loss_train: null model");
    return S3D_ERR_BAD_ARG;
  }
  return vgg_loss_train_bwd(m, N, S, gout_dev, saved_dev, saved_bytes, grad_a_dev, static_cast<cudaStream_t>(stream));
}

size_t s3d_gt_encoder_workspace_bytes(int32_t B, int32_t K, int32_t S) {
  if (B <= 0 || K <= 0 || S <= 0) return 0;
  return gt_encoder_workspace_bytes(B * K, S);
}
int s3d_gt_encoder_fwd(const s3d_model* m, const float* img_slices_dev, int32_t B, int32_t S, void* planes_dev,
                       float* const* taps_nchw_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
  return gt_encoder_fwd(m, img_slices_dev, B, S, planes_dev, taps_nchw_dev, workspace_dev, workspace_bytes,
                        static_cast<cudaStream_t>(stream));
}
size_t s3d_gt_decoder_workspace_bytes(int64_t n, int32_t precision) { return gt_decoder_workspace_bytes(n, precision); }
int s3d_gt_decoder_fwd(const s3d_model* m, const void* planes_dev, int32_t S, float* qry_dev, int32_t B, int64_t n_per_image,
                       const float* T_dev, const float* rot_dev, int32_t flip_in_place, float out_scale, float* out_dev,
                       int32_t precision, void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (B < 1 || n_per_image < 0) {
    set_error("gt_decoder: bad batch");
    return S3D_ERR_BAD_ARG;
  }
  const int64_t n = (int64_t)B * n_per_image;
  if (n == 0) return S3D_OK;
  if (!qry_dev) {
    set_error("gt_decoder: null query pointer");
    return S3D_ERR_BAD_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  QueryCtx q{};
  q.qry = qry_dev;
  q.T = T_dev;
  q.rot = rot_dev;
  if (B > 1) {
    q.per_img = n_per_image;
    q.plane_stride = plane_offset_floats(m ? m->K : 0, S, 5);
  }
  S3D_TRY(check_decoder(m, planes_dev, S, T_dev, n, out_dev, precision, nullptr, workspace_dev, workspace_bytes, true));
  if (!rot_dev && flip_in_place) {
    k_flip_yz<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(qry_dev, n);
    S3D_LAUNCH_CHECK();
    q.preflipped = 1;
  }
  return gt_decoder_fwd(m, planes_dev, S, q, n, out_scale, out_dev, precision, workspace_dev, workspace_bytes, st);
}

size_t s3d_preprocess_workspace_bytes(int32_t N, int32_t H, int32_t S) {
  if (N < 1 || H < 1 || S < 1) return 0;
  return preprocess_workspace_bytes(N, H, S);
}
int s3d_preprocess_rgba(const uint8_t* rgba_dev, int32_t N, int32_t H, int32_t W, int32_t S, int32_t white_bg,
                        const int32_t* bounds_h_dev, const int32_t* kk_h_dev, int32_t ksize_h, const int32_t* bounds_v_dev,
                        const int32_t* kk_v_dev, int32_t ksize_v, float* out_dev, void* workspace_dev, size_t workspace_bytes,
                        void* stream) {
  return preprocess_rgba(rgba_dev, N, H, W, S, white_bg, bounds_h_dev, kk_h_dev, ksize_h, bounds_v_dev, kk_v_dev, ksize_v,
                         out_dev, workspace_dev, workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t s3d_sparse_scratch_bytes(int32_t resolution0, int32_t depth, int64_t capacity) {
  if (resolution0 < 1 || depth < 0 || depth > 15 || capacity < 1) return 0;
  const size_t nblk = mise_query_blocks(resolution0 << depth);
  return (nblk + 64) * sizeof(int) + (size_t)capacity * (sizeof(int) + 4 * sizeof(float)) + 256;
}

int s3d_sparse_rounds(const s3d_model* m, const void* planes_dev, int32_t S, const float* T_dev, double box_size,
                      float out_scale, int32_t resolution0, int32_t depth, double threshold, double* value_dev,
                      uint8_t* known_dev, int8_t* cell_level_dev, uint8_t* exists_dev, int32_t* flags_dev, void* scratch_dev,
                      int64_t capacity, int32_t* counts_dev, int32_t n_rounds, int32_t precision, void* workspace_dev,
                      size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!value_dev || !known_dev || !cell_level_dev || !exists_dev || !flags_dev || !scratch_dev || !counts_dev ||
      n_rounds < 1 || capacity < 1 || capacity > 0x7fffffff || resolution0 < 1 || depth < 0 || depth > 15) {
    set_error("sparse_rounds: bad argument");
    return S3D_ERR_BAD_ARG;
  }
  if (precision == S3D_PREC_FP32) {
    set_error("sparse_rounds: needs a tensor-core precision mode (the fp32 decoder keeps the host loop)");
    return S3D_ERR_UNSUPPORTED;
  }
  float* out_probe = reinterpret_cast<float*>(scratch_dev);
  S3D_TRY(check_decoder(m, planes_dev, S, T_dev, capacity, out_probe, precision, nullptr, workspace_dev, workspace_bytes));
  const int R = resolution0 << depth;
  const size_t nblk = mise_query_blocks(R);
  int* blk = static_cast<int*>(scratch_dev);
  int* pt_idx = blk + nblk + 64;
  float* pts = reinterpret_cast<float*>(pt_idx + capacity);
  float* vals = pts + 3 * (size_t)capacity;
  int* live = counts_dev + n_rounds;
  int* overflow = counts_dev + n_rounds + 1;
  for (int r = 0; r < n_rounds; ++r) {
    S3D_TRY(mise_query_device(R, box_size, exists_dev, known_dev, blk, live, (int)capacity, counts_dev + r, overflow, pt_idx,
                              pts, st));
    QueryCtx q{};
    q.qry = pts;  // test mode: y,z negated on the fly (our own buffer: nothing to write back)
    q.T = T_dev;
    S3D_TRY(decoder_tc(m, static_cast<const float*>(planes_dev), S, q, capacity, out_scale, vals, precision, workspace_dev,
                       workspace_bytes, st, live));
    S3D_TRY(mise_apply_device(live, pt_idx, vals, value_dev, known_dev, st));
    S3D_TRY(mise_subdivide(resolution0, depth, threshold, value_dev, known_dev, reinterpret_cast<signed char*>(cell_level_dev),
                           exists_dev, flags_dev, st));
  }
  return S3D_OK;
}

int s3d_debug_set_decoder_flags(int32_t flags) {
  decoder_tc_set_debug(flags);
  return S3D_OK;
}

int s3d_debug_set_encoder(s3d_model* m, int32_t simt) {
  if (!m) return S3D_ERR_BAD_ARG;
  m->enc_simt = simt ? 1 : 0;
  return S3D_OK;
}

size_t s3d_scan_scratch_bytes(int64_t n) { return n > 0 ? scan_scratch_bytes(n) : 0; }
int s3d_exclusive_scan(const int32_t* in_dev, int64_t n, int64_t* out_dev, int64_t* total_dev, void* scratch_dev, void* stream) {
  if (!in_dev || !out_dev || !total_dev || !scratch_dev || n < 0) {
    set_error("exclusive_scan: bad argument");
    return S3D_ERR_BAD_ARG;
  }
  return exclusive_scan_i32_i64(in_dev, n, reinterpret_cast<long long*>(out_dev), reinterpret_cast<long long*>(total_dev),
                                scratch_dev, static_cast<cudaStream_t>(stream));
}

int s3d_debug_profile(int64_t* out32, int32_t reset) { return debug_profile(reinterpret_cast<long long*>(out32), reset); }

int s3d_selftest_umma(int32_t mode, int32_t passes, const float* a_dev, const float* w_dev, float* d_dev,
                      void* stream) {
  return umma_selftest(mode, passes, a_dev, w_dev, d_dev, static_cast<cudaStream_t>(stream));
}

int s3d_decoder_debug_tokens(const s3d_model* m, const void* planes_dev, int32_t S, const float* qry_dev, int64_t n,
                             const float* T_dev, const float* rot_dev, float* out_dev, float* tokens_dev,
                             void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (!qry_dev || !tokens_dev) {
    set_error("decoder_debug: null pointer");
    return S3D_ERR_BAD_ARG;
  }
  QueryCtx q{};
  q.qry = qry_dev;
  q.T = T_dev;
  q.rot = rot_dev;
  return run_decoder(m, planes_dev, S, q, n, 1.f, out_dev, S3D_PREC_FP32, tokens_dev, workspace_dev, workspace_bytes,
                     static_cast<cudaStream_t>(stream));
}

}  // extern "C"
