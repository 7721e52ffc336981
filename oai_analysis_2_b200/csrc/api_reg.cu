// C-ABI entry points of the registration / warping stage.
#include "../../include/oai_b200.h"
#include "api_common.h"
#include "reg_kernels.cuh"

using namespace oai;

namespace {
Affine3 affine_from(const double* a) {
  Affine3 r;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) r.m[3 * i + j] = a[4 * i + j];
    r.t[i] = a[4 * i + 3];
  }
  return r;
}
}  // namespace

extern "C" size_t oai_reg_conv3_workspace(int cin, int cout, const int* in_dims, int N, int stride, int leaky_in) {
  if (!in_dims || cin <= 0 || cout <= 0 || N <= 0 || stride < 1) return 0;
  Conv3Params p{};
  p.cin = cin; p.cout = cout; p.N = N; p.stride = stride; p.leaky_in = leaky_in;
  p.Di = in_dims[0]; p.Hi = in_dims[1]; p.Wi = in_dims[2];
  p.Do = (p.Di - 1) / stride + 1; p.Ho = (p.Hi - 1) / stride + 1; p.Wo = (p.Wi - 1) / stride + 1;
  return conv3_splitk_bytes(p);
}

extern "C" int oai_reg_conv3(const float* in, long long in_nstride, long long in_cstride, int cin, const int* in_dims,
                             const float* w, const float* bias, float* out, long long out_nstride,
                             long long out_cstride, int cout, int cout_pad, int N, int stride, int leaky_in,
                             int residual, float out_scale, void* workspace, size_t workspace_bytes, void* stream) {
  OAI_REQUIRE(in && in_dims && w && bias && out, "reg_conv3: null pointer");
  OAI_REQUIRE(stride == 1 || stride == 2, "reg_conv3: stride %d unsupported", stride);
  OAI_REQUIRE(!residual || (stride == 2 && cout >= cin), "reg_conv3: residual needs stride 2 and cout >= cin");
  Conv3Params p{};
  p.in = in; p.in_nstride = in_nstride; p.in_cstride = in_cstride; p.cin = cin;
  p.Di = in_dims[0]; p.Hi = in_dims[1]; p.Wi = in_dims[2];
  p.w = w; p.bias = bias; p.out = out; p.out_nstride = out_nstride; p.out_cstride = out_cstride;
  p.cout = cout; p.cout_pad = cout_pad;
  p.Do = (p.Di - 1) / stride + 1; p.Ho = (p.Hi - 1) / stride + 1; p.Wo = (p.Wi - 1) / stride + 1;
  p.N = N; p.stride = stride; p.leaky_in = leaky_in; p.residual = residual; p.out_scale = out_scale;
  p.splitk_ws = static_cast<float*>(workspace); p.splitk_bytes = workspace ? workspace_bytes : 0;
  return conv3_launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" int oai_reg_convt4(const float* in, long long in_nstride, long long in_cstride, int cin, const int* in_dims,
                              const float* w, const float* bias, const float* bn_scale, const float* bn_shift,
                              float* out, long long out_nstride, long long out_cstride, int cout, const int* out_dims,
                              int N, void* stream) {
  OAI_REQUIRE(in && in_dims && w && bias && bn_scale && bn_shift && out && out_dims, "reg_convt4: null pointer");
  OAI_REQUIRE(cout <= cin, "reg_convt4: the residual keeps the first cout of cin channels (cout=%d cin=%d)", cout, cin);
  for (int a = 0; a < 3; ++a)
    OAI_REQUIRE(out_dims[a] >= 1 && out_dims[a] <= 2 * in_dims[a], "reg_convt4: output dim %d exceeds 2x input", a);
  ConvT4Params p{};
  p.in = in; p.in_nstride = in_nstride; p.in_cstride = in_cstride; p.cin = cin;
  p.Di = in_dims[0]; p.Hi = in_dims[1]; p.Wi = in_dims[2];
  p.w = w; p.bias = bias; p.bn_scale = bn_scale; p.bn_shift = bn_shift;
  p.out = out; p.out_nstride = out_nstride; p.out_cstride = out_cstride; p.cout = cout;
  p.Do = out_dims[0]; p.Ho = out_dims[1]; p.Wo = out_dims[2]; p.N = N;
  p.wpk = nullptr; p.wexp = 0; p.xsplit = nullptr; p.xsplit_bytes = 0; p.debug = 0;
  return convt4_launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" int oai_reg_pack_convt4(const float* w, int cin, int cout, int wexp, void* wpk, void* stream) {
  OAI_REQUIRE(w && wpk, "reg_pack_convt4: null pointer");
  OAI_REQUIRE(cin > 0 && cout > 0 && cin % 16 == 0 && cout % 16 == 0,
              "reg_pack_convt4: cin and cout must be multiples of 16 (got %d, %d)", cin, cout);
  OAI_REQUIRE(wexp >= -14 && wexp <= 30, "reg_pack_convt4: scale exponent %d out of range", wexp);
  return reg_pack_convt4_launch(w, cin, cout, wexp, static_cast<uint4*>(wpk), static_cast<cudaStream_t>(stream));
}

extern "C" size_t oai_reg_convt4_mma_workspace(int cin, int cout, const int* in_dims, int N) {
  if (!in_dims || cin <= 0 || cout <= 0 || N <= 0) return 0;
  ConvT4Params p{};
  p.cin = cin; p.cout = cout; p.N = N; p.Di = in_dims[0]; p.Hi = in_dims[1]; p.Wi = in_dims[2];
  const size_t deep = convt4_splitk_bytes(p);   // levels narrower than 12 points: split-K partial sums
  return deep ? deep : static_cast<size_t>(N) * cin * in_dims[0] * in_dims[1] * in_dims[2] * 4;
}

extern "C" int oai_reg_convt4_mma(const float* in, long long in_nstride, long long in_cstride, int cin,
                                  const int* in_dims, const float* w, const void* wpk, int wexp, const float* bias,
                                  const float* bn_scale, const float* bn_shift, float* out, long long out_nstride,
                                  long long out_cstride, int cout, const int* out_dims, int N, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  OAI_REQUIRE(in && in_dims && w && wpk && bias && bn_scale && bn_shift && out && out_dims && workspace,
              "reg_convt4_mma: null pointer");
  OAI_REQUIRE(workspace_bytes >= oai_reg_convt4_mma_workspace(cin, cout, in_dims, N) &&
                  (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
              "reg_convt4_mma: workspace of %zu bytes, 16-byte aligned, required",
              oai_reg_convt4_mma_workspace(cin, cout, in_dims, N));
  OAI_REQUIRE(cout <= cin, "reg_convt4_mma: the residual keeps the first cout of cin channels (cout=%d cin=%d)", cout,
              cin);
  OAI_REQUIRE(cin % 16 == 0 && cout % 16 == 0, "reg_convt4_mma: cin and cout must be multiples of 16");
  OAI_REQUIRE((reinterpret_cast<uintptr_t>(wpk) & 15) == 0, "reg_convt4_mma: packed weights must be 16-byte aligned");
  for (int a = 0; a < 3; ++a)
    OAI_REQUIRE(out_dims[a] >= 1 && out_dims[a] <= 2 * in_dims[a], "reg_convt4_mma: output dim %d exceeds 2x input", a);
  ConvT4Params p{};
  p.in = in; p.in_nstride = in_nstride; p.in_cstride = in_cstride; p.cin = cin;
  p.Di = in_dims[0]; p.Hi = in_dims[1]; p.Wi = in_dims[2];
  p.w = w; p.bias = bias; p.bn_scale = bn_scale; p.bn_shift = bn_shift;
  p.out = out; p.out_nstride = out_nstride; p.out_cstride = out_cstride; p.cout = cout;
  p.Do = out_dims[0]; p.Ho = out_dims[1]; p.Wo = out_dims[2]; p.N = N;
  p.wpk = static_cast<const uint4*>(wpk); p.wexp = wexp; p.xsplit = static_cast<uint32_t*>(workspace);
  p.xsplit_bytes = workspace_bytes; p.debug = 0;
  return convt4_launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" size_t oai_reg_conv3_umma_wbytes(int cin, int cout) { return conv3_umma_wbytes(cin, cout); }

extern "C" size_t oai_reg_conv3_umma_workspace(int cin, const int* in_dims, int N) {
  if (!in_dims || cin <= 0 || N <= 0) return 0;
  Conv3Params p{};
  p.cin = cin; p.N = N; p.Di = in_dims[0]; p.Hi = in_dims[1]; p.Wi = in_dims[2];
  return conv3_umma_workspace(p);
}

extern "C" int oai_reg_pack_conv3_umma(const float* w, int cin, int cout, int cout_pad, int wexp, void* dst,
                                       void* stream) {
  OAI_REQUIRE(w && dst, "reg_pack_conv3_umma: null pointer");
  OAI_REQUIRE(conv3_umma_wbytes(cin, cout) > 0 && cout_pad >= cout,
              "reg_pack_conv3_umma: cin must be a multiple of 16 and cout 32 or a multiple of 64 (got %d, %d)", cin, cout);
  OAI_REQUIRE(wexp >= -14 && wexp <= 30, "reg_pack_conv3_umma: scale exponent %d out of range", wexp);
  OAI_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 15) == 0, "reg_pack_conv3_umma: dst must be 16-byte aligned");
  return reg_pack_conv3_umma_launch(w, cin, cout, cout_pad, wexp, dst, static_cast<cudaStream_t>(stream));
}

extern "C" int oai_reg_conv3_umma(const float* in, long long in_nstride, long long in_cstride, int cin,
                                  const int* in_dims, const void* wumma, int wexp, const float* bias, float* out,
                                  long long out_nstride, long long out_cstride, int cout, int N, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  OAI_REQUIRE(in && in_dims && wumma && bias && out && workspace, "reg_conv3_umma: null pointer");
  OAI_REQUIRE((reinterpret_cast<uintptr_t>(wumma) & 15) == 0, "reg_conv3_umma: weight blocks must be 16-byte aligned");
  Conv3Params p{};
  p.in = in; p.in_nstride = in_nstride; p.in_cstride = in_cstride; p.cin = cin;
  p.Di = in_dims[0]; p.Hi = in_dims[1]; p.Wi = in_dims[2];
  p.bias = bias; p.out = out; p.out_nstride = out_nstride; p.out_cstride = out_cstride;
  p.cout = cout; p.cout_pad = cout;
  p.Do = (p.Di + 1) / 2; p.Ho = (p.Hi + 1) / 2; p.Wo = (p.Wi + 1) / 2;
  p.N = N; p.stride = 2; p.leaky_in = 1; p.residual = 1; p.out_scale = 1.0f;
  p.wumma = wumma; p.wexp = wexp; p.xsplit = workspace; p.xsplit_bytes = workspace_bytes;
  OAI_REQUIRE(conv3_umma_eligible(p),
              "reg_conv3_umma: needs cin %% 16 == 0, cout 32 or a multiple of 64 (<= 512), cout >= cin and an output "
              "lattice of at least 8 x 8 (got cin=%d cout=%d input %d x %d x %d)", cin, cout, p.Di, p.Hi, p.Wi);
  OAI_REQUIRE(workspace_bytes >= conv3_umma_workspace(p) && (reinterpret_cast<uintptr_t>(workspace) & 127) == 0,
              "reg_conv3_umma: workspace of %zu bytes, 128-byte aligned, required", conv3_umma_workspace(p));
  return conv3_umma_launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" size_t oai_reg_convt4_umma_wbytes(int cin, int cout) { return convt4_umma_wbytes(cin, cout); }

extern "C" int oai_reg_pack_convt4_umma(const float* w, int cin, int cout, int wexp, void* dst, void* stream) {
  OAI_REQUIRE(w && dst, "reg_pack_convt4_umma: null pointer");
  OAI_REQUIRE(cin > 0 && cin % 16 == 0 && (cout == 16 || cout == 32 || cout == 64 || cout == 128),
              "reg_pack_convt4_umma: cin must be a multiple of 16 and cout one of 16, 32, 64, 128 (got %d, %d)", cin,
              cout);
  OAI_REQUIRE(wexp >= -14 && wexp <= 30, "reg_pack_convt4_umma: scale exponent %d out of range", wexp);
  OAI_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 15) == 0, "reg_pack_convt4_umma: dst must be 16-byte aligned");
  return reg_pack_convt4_umma_launch(w, cin, cout, wexp, dst, static_cast<cudaStream_t>(stream));
}

extern "C" int oai_reg_convt4_umma(const float* in, long long in_nstride, long long in_cstride, int cin,
                                   const int* in_dims, const void* wumma, int wexp, const float* bias,
                                   const float* bn_scale, const float* bn_shift, float* out, long long out_nstride,
                                   long long out_cstride, int cout, const int* out_dims, int N, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  OAI_REQUIRE(in && in_dims && wumma && bias && bn_scale && bn_shift && out && out_dims && workspace,
              "reg_convt4_umma: null pointer");
  OAI_REQUIRE((reinterpret_cast<uintptr_t>(wumma) & 15) == 0, "reg_convt4_umma: weight blocks must be 16-byte aligned");
  for (int a = 0; a < 3; ++a)
    OAI_REQUIRE(out_dims[a] >= 1 && out_dims[a] <= 2 * in_dims[a], "reg_convt4_umma: output dim %d exceeds 2x input", a);
  ConvT4Params p{};
  p.in = in; p.in_nstride = in_nstride; p.in_cstride = in_cstride; p.cin = cin;
  p.Di = in_dims[0]; p.Hi = in_dims[1]; p.Wi = in_dims[2];
  p.bias = bias; p.bn_scale = bn_scale; p.bn_shift = bn_shift;
  p.out = out; p.out_nstride = out_nstride; p.out_cstride = out_cstride; p.cout = cout;
  p.Do = out_dims[0]; p.Ho = out_dims[1]; p.Wo = out_dims[2]; p.N = N;
  p.wexp = wexp; p.wumma = wumma; p.xsplit = static_cast<uint32_t*>(workspace); p.xsplit_bytes = workspace_bytes;
  OAI_REQUIRE(convt4_umma_eligible(p),
              "reg_convt4_umma: needs cout in {16, 32, 64, 128}, cin %% 16 == 0, cout <= cin and a lattice of at least 8 x 8 "
              "(got cin=%d cout=%d dims %d x %d x %d)", cin, cout, p.Di, p.Hi, p.Wi);
  return convt4_umma_launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" int oai_compose(const int* grid_dims, int nfields, const float* const* fields, const int* field_dims,
                           int shortcut_first, const float* img, const int* img_dims, float* phi_out, float* img_out,
                           void* stream) {
  OAI_REQUIRE(grid_dims && nfields >= 0 && nfields <= 4, "compose: 0..4 fields supported (got %d)", nfields);
  OAI_REQUIRE(phi_out || img_out, "compose: nothing to write");
  OAI_REQUIRE(!img_out || (img && img_dims), "compose: img_out needs img");
  ChainParams p;
  p.D = grid_dims[0]; p.H = grid_dims[1]; p.W = grid_dims[2];
  OAI_REQUIRE(p.D > 1 && p.H > 1 && p.W > 1, "compose: grid axes must have at least 2 samples");
  p.nfields = nfields;
  for (int f = 0; f < 4; ++f) {
    p.u[f] = f < nfields ? fields[f] : nullptr;
    p.ud[f] = f < nfields ? field_dims[3 * f] : 0;
    p.uh[f] = f < nfields ? field_dims[3 * f + 1] : 0;
    p.uw[f] = f < nfields ? field_dims[3 * f + 2] : 0;
    OAI_REQUIRE(f >= nfields || p.u[f], "compose: null field %d", f);
  }
  p.shortcut_first = shortcut_first;
  if (shortcut_first)
    OAI_REQUIRE(nfields > 0 && p.ud[0] == p.D && p.uh[0] == p.H && p.uw[0] == p.W,
                "compose: the identity shortcut needs the first field on the grid itself");
  p.img = img_out ? img : nullptr;
  p.id = img_dims ? img_dims[0] : 0; p.ih = img_dims ? img_dims[1] : 0; p.iw = img_dims ? img_dims[2] : 0;
  p.phi_out = phi_out; p.img_out = img_out;
  return chain_launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" int oai_resize_trilinear(const float* in, const int* in_dims, float* out, const int* out_dims,
                                    void* stream) {
  OAI_REQUIRE(in && out && in_dims && out_dims, "resize: null pointer");
  return resize_trilinear_launch(in, in_dims[0], in_dims[1], in_dims[2], out, out_dims[0], out_dims[1], out_dims[2],
                                 static_cast<cudaStream_t>(stream));
}

extern "C" int oai_avgpool3d_2_ceil(const float* in, int C, const int* in_dims, float* out, void* stream) {
  OAI_REQUIRE(in && out && in_dims, "avgpool: null pointer");
  return avgpool2_ceil_launch(in, C, in_dims[0], in_dims[1], in_dims[2], out, static_cast<cudaStream_t>(stream));
}

extern "C" int oai_displacement_field(const float* phi, const int* dims, float* disp, void* stream) {
  OAI_REQUIRE(phi && disp && dims, "displacement_field: null pointer");
  return disp_field_launch(phi, dims[0], dims[1], dims[2], disp, static_cast<cudaStream_t>(stream));
}

extern "C" int oai_warp_volume(const float* src, int C, const int* src_dims, const float* disp, const int* field_dims,
                               const double* out_index_to_net, const double* net_to_src_index, float* out,
                               const int* out_dims, float default_value, void* stream) {
  OAI_REQUIRE(src && src_dims && disp && field_dims && out_index_to_net && net_to_src_index && out && out_dims,
              "warp_volume: null pointer");
  WarpVolumeParams p;
  p.src = src; p.C = C; p.SD = src_dims[0]; p.SH = src_dims[1]; p.SW = src_dims[2];
  p.disp = disp; p.FD = field_dims[0]; p.FH = field_dims[1]; p.FW = field_dims[2];
  p.out_index_to_net = affine_from(out_index_to_net);
  p.net_to_src_index = affine_from(net_to_src_index);
  p.out = out; p.OD = out_dims[0]; p.OH = out_dims[1]; p.OW = out_dims[2];
  p.default_value = default_value;
  return warp_volume_launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" int oai_warp_points(const double* pts, long long n, const float* disp, const int* field_dims,
                               const double* phys_to_net, const double* net_to_phys, double* out, void* stream) {
  OAI_REQUIRE(disp && field_dims && phys_to_net && net_to_phys && (n == 0 || (pts && out)),
              "warp_points: null pointer");
  if (n == 0) return 0;
  WarpPointsParams p;
  p.pts = pts; p.out = out; p.n = n; p.disp = disp;
  p.FD = field_dims[0]; p.FH = field_dims[1]; p.FW = field_dims[2];
  p.phys_to_net = affine_from(phys_to_net);
  p.net_to_phys = affine_from(net_to_phys);
  return warp_points_launch(p, static_cast<cudaStream_t>(stream));
}
