// Ref-NeRF per-sample stage between the spatial trunk and the directional MLP, and its adjoint.
//
// Forward replaces (reference file:line): normals / normals_pred = -l2_normalize(.)
// models.py:488-499 + ref_utils.l2_normalize ref_utils.py:40-42; roughness models.py:520-523;
// ref_utils.reflect ref_utils.py:22-37 (models.py:545); the integrated directional encoding
// ref_utils.generate_ide_fn ref_utils.py:98-159 or coord.pos_enc for plain view directions;
// n.v models.py:560-563.  It writes the bf16 direction-encoding slab of the view-MLP input.
// Backward fuses the adjoint of all of the above with train_utils.orientation_loss
// train_utils.py:162-178 and train_utils.predicted_normal_loss :181-197.
// One thread per sample: HBM-bound elementwise work.
#include <algorithm>

#include "common.cuh"

namespace mnrf {

constexpr int kIdeMax = 36;     // (m,l) pairs at deg_view = 5
constexpr int kZMax = 17;       // z^0 .. z^16

struct IdeTab {                 // staged in shared memory
  float mat[kZMax * kIdeMax];   // [k][i]
  float sigma[kIdeMax];
  int m[kIdeMax];
  int n, zdeg;
};

__device__ __forceinline__ void load_tab(IdeTab* s, const float* __restrict__ mat, const int* __restrict__ ml,
                                         int n, int zdeg) {
  for (int i = threadIdx.x; i < zdeg * n; i += blockDim.x) s->mat[i] = mat[i];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    s->m[i] = ml[i];
    int l = ml[n + i];
    s->sigma[i] = 0.5f * (float)l * (float)(l + 1);
  }
  if (threadIdx.x == 0) { s->n = n; s->zdeg = zdeg; }
  __syncthreads();
}

// -x / sqrt(max(|x|^2, eps))
__device__ __forceinline__ void neg_normalize(const float g[3], float out[3], float& s, bool& clamped) {
  float sq = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
  clamped = !(sq > kEps);
  s = sqrtf(fmaxf(sq, kEps));
  out[0] = -g[0] / s; out[1] = -g[1] / s; out[2] = -g[2] / s;
}
// adjoint of neg_normalize: given a = dL/dout, returns dL/dg
__device__ __forceinline__ void neg_normalize_bwd(const float out[3], float s, bool clamped, const float a[3],
                                                  float dg[3]) {
  if (clamped) { dg[0] = -a[0] / s; dg[1] = -a[1] / s; dg[2] = -a[2] / s; return; }
  // out = -ghat:  d out/d g = -(I - ghat ghat^T)/s = -(I - out out^T)/s
  float dot = out[0] * a[0] + out[1] * a[1] + out[2] * a[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) dg[i] = -(a[i] - out[i] * dot) / s;
}

struct RefDesc {
  int64_t M;
  int S;                       // samples per ray (viewdirs are per ray)
  int use_pred_normals, use_density_normals, use_reflections, use_ide, use_n_dot_v, use_roughness;
  int deg_view;
  float roughness_bias;
  int ld, col0, col_end;       // bf16 slab [M, ld], columns [col0, col_end)
};

struct RefLoss {
  float orient_mult, prednorm_mult;   // already divided by the number of rays
  int orient_on_pred;                 // orientation_loss_target == 'normals_pred'
};

__global__ void __launch_bounds__(128)
refdir_fwd_kernel(RefDesc d, const float* __restrict__ ide_mat, const int* __restrict__ ide_ml, int ide_n,
                  const float* __restrict__ grad_pred, const float* __restrict__ raw_rough,
                  const float* __restrict__ raw_grad_density /* [3, M] */, const float* __restrict__ viewdirs,
                  float* __restrict__ normals_pred, float* __restrict__ normals, float* __restrict__ roughness,
                  __nv_bfloat16* __restrict__ slab, RefLoss L, float* __restrict__ extra_dw) {
  __shared__ IdeTab tab;
  if (d.use_ide) load_tab(&tab, ide_mat, ide_ml, ide_n, (1 << (d.deg_view - 1)) + 1);
  for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < d.M; m += (int64_t)gridDim.x * blockDim.x) {
    const int ray = (int)(m / d.S);
    const float v[3] = {viewdirs[ray * 3], viewdirs[ray * 3 + 1], viewdirs[ray * 3 + 2]};
    float np_[3] = {0.f, 0.f, 0.f}, nd[3] = {0.f, 0.f, 0.f}, s;
    bool cl;
    if (d.use_pred_normals) {
      const float g[3] = {grad_pred[m * 3], grad_pred[m * 3 + 1], grad_pred[m * 3 + 2]};
      neg_normalize(g, np_, s, cl);
      normals_pred[m * 3] = np_[0]; normals_pred[m * 3 + 1] = np_[1]; normals_pred[m * 3 + 2] = np_[2];
    }
    if (d.use_density_normals) {
      const float g[3] = {raw_grad_density[m], raw_grad_density[d.M + m], raw_grad_density[2 * d.M + m]};
      neg_normalize(g, nd, s, cl);
      normals[m * 3] = nd[0]; normals[m * 3 + 1] = nd[1]; normals[m * 3 + 2] = nd[2];
    }
    const float* n = d.use_pred_normals ? np_ : nd;
    float kappa = 0.f;
    if (d.use_roughness) {
      kappa = softplus_f(raw_rough[m] + d.roughness_bias);
      roughness[m] = kappa;
    }
    if (extra_dw) {
      // d(orientation + predicted-normal loss)/d(weight of this sample): pure forward quantities
      float dw = 0.f;
      if (L.orient_mult > 0.f) {
        const float* no = L.orient_on_pred ? np_ : nd;
        float pm = fminf(0.f, -(no[0] * v[0] + no[1] * v[1] + no[2] * v[2]));
        dw += L.orient_mult * pm * pm;
      }
      if (L.prednorm_mult > 0.f) dw += L.prednorm_mult * (1.f - (nd[0] * np_[0] + nd[1] * np_[1] + nd[2] * np_[2]));
      extra_dw[m] = dw;
    }
    const float ndv = n[0] * v[0] + n[1] * v[1] + n[2] * v[2];
    float dir[3] = {v[0], v[1], v[2]};
    if (d.use_reflections) {
      // reflect(-v, n) = 2 (n . -v) n + v
#pragma unroll
      for (int i = 0; i < 3; ++i) dir[i] = v[i] - 2.f * ndv * n[i];
    }
    __nv_bfloat16* out = slab + m * (int64_t)d.ld + d.col0;
    int c = 0;
    if (d.use_ide) {
      float zp[kZMax];
      zp[0] = 1.f;
      for (int k = 1; k < tab.zdeg; ++k) zp[k] = zp[k - 1] * dir[2];
      // (x + iy)^m by repeated multiplication; pairs are listed with m increasing inside each l
      for (int i = 0; i < tab.n; ++i) {
        float P = 0.f;
        for (int k = 0; k < tab.zdeg; ++k) P += zp[k] * tab.mat[k * tab.n + i];
        float cr = 1.f, ci = 0.f;
        for (int q = 0; q < tab.m[i]; ++q) { float t = cr * dir[0] - ci * dir[1]; ci = cr * dir[1] + ci * dir[0]; cr = t; }
        float A = expf(-tab.sigma[i] * kappa);
        out[i] = __float2bfloat16(cr * P * A);
        out[tab.n + i] = __float2bfloat16(ci * P * A);
      }
      c = 2 * tab.n;
    } else {
      // coord.pos_enc(dir, 0, deg_view, append_identity=True)
      out[0] = __float2bfloat16(dir[0]); out[1] = __float2bfloat16(dir[1]); out[2] = __float2bfloat16(dir[2]);
      for (int half = 0; half < 2; ++half)
        for (int l = 0; l < d.deg_view; ++l)
          for (int ch = 0; ch < 3; ++ch) {
            float x = dir[ch] * exp2f((float)l);
            out[3 + half * 3 * d.deg_view + l * 3 + ch] = __float2bfloat16(sinf(half ? x + 1.57079637050628662109375f : x));
          }
      c = 3 + 6 * d.deg_view;
    }
    if (d.use_n_dot_v) out[c++] = __float2bfloat16(ndv);
    for (; d.col0 + c < d.col_end; ++c) out[c] = __float2bfloat16(0.f);
  }
}

__global__ void __launch_bounds__(128)
refdir_bwd_kernel(RefDesc d, RefLoss L, const float* __restrict__ ide_mat, const int* __restrict__ ide_ml, int ide_n,
                  const float* __restrict__ grad_pred, const float* __restrict__ raw_rough,
                  const float* __restrict__ raw_grad_density, const float* __restrict__ viewdirs,
                  const float* __restrict__ weights, __nv_bfloat16* __restrict__ d_slab, int ld_dslab,
                  const float* __restrict__ d_raw_density, const float* __restrict__ d_raw_diffuse,
                  const float* __restrict__ d_raw_tint,
                  float* __restrict__ d_grad_pred, float* __restrict__ d_raw_rough,
                  float* __restrict__ d_raw_grad_density /* [3, M] */,
                  float* __restrict__ stats /* [4]=orientation, [5]=pred normals */) {
  __shared__ IdeTab tab;
  if (d.use_ide) load_tab(&tab, ide_mat, ide_ml, ide_n, (1 << (d.deg_view - 1)) + 1);
  float st_or = 0.f, st_pn = 0.f;
  for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < d.M; m += (int64_t)gridDim.x * blockDim.x) {
    const int ray = (int)(m / d.S);
    const float v[3] = {viewdirs[ray * 3], viewdirs[ray * 3 + 1], viewdirs[ray * 3 + 2]};
    float np_[3] = {0.f, 0.f, 0.f}, nd[3] = {0.f, 0.f, 0.f}, s_p = 1.f, s_d = 1.f;
    bool cl_p = false, cl_d = false;
    if (d.use_pred_normals) {
      const float g[3] = {grad_pred[m * 3], grad_pred[m * 3 + 1], grad_pred[m * 3 + 2]};
      neg_normalize(g, np_, s_p, cl_p);
    }
    if (d.use_density_normals) {
      const float g[3] = {raw_grad_density[m], raw_grad_density[d.M + m], raw_grad_density[2 * d.M + m]};
      neg_normalize(g, nd, s_d, cl_d);
    }
    const float* n = d.use_pred_normals ? np_ : nd;
    float kappa = 0.f, rin = 0.f;
    if (d.use_roughness) { rin = raw_rough[m] + d.roughness_bias; kappa = softplus_f(rin); }
    const float ndv = n[0] * v[0] + n[1] * v[1] + n[2] * v[2];
    float dir[3] = {v[0], v[1], v[2]};
    if (d.use_reflections) {
#pragma unroll
      for (int i = 0; i < 3; ++i) dir[i] = v[i] - 2.f * ndv * n[i];
    }
    // ---- adjoint of the direction encoding: u = dL/d dir, dkappa
    const __nv_bfloat16* gin = d_slab + m * (int64_t)ld_dslab + d.col0;
    float u[3] = {0.f, 0.f, 0.f}, dkappa = 0.f;
    int c = 0;
    if (d.use_ide) {
      float zp[kZMax];
      zp[0] = 1.f;
      for (int k = 1; k < tab.zdeg; ++k) zp[k] = zp[k - 1] * dir[2];
      for (int i = 0; i < tab.n; ++i) {
        float P = 0.f, dP = 0.f;
        for (int k = 0; k < tab.zdeg; ++k) {
          float co = tab.mat[k * tab.n + i];
          P += zp[k] * co;
          if (k > 0) dP += (float)k * zp[k - 1] * co;
        }
        const int mi = tab.m[i];
        float cr = 1.f, ci = 0.f, er = 0.f, ei = 0.f;      // c = (x+iy)^m, e = m (x+iy)^(m-1)
        for (int q = 0; q < mi; ++q) {
          if (q == mi - 1) { er = (float)mi * cr; ei = (float)mi * ci; }
          float t = cr * dir[0] - ci * dir[1]; ci = cr * dir[1] + ci * dir[0]; cr = t;
        }
        const float A = expf(-tab.sigma[i] * kappa);
        const float gr = __bfloat162float(gin[i]), gi = __bfloat162float(gin[tab.n + i]);
        dkappa += -tab.sigma[i] * A * P * (gr * cr + gi * ci);
        u[2] += A * dP * (gr * cr + gi * ci);
        u[0] += A * P * (gr * er + gi * ei);
        u[1] += A * P * (-gr * ei + gi * er);
      }
      c = 2 * tab.n;
    } else {
      u[0] = __bfloat162float(gin[0]); u[1] = __bfloat162float(gin[1]); u[2] = __bfloat162float(gin[2]);
      for (int half = 0; half < 2; ++half)
        for (int l = 0; l < d.deg_view; ++l)
          for (int ch = 0; ch < 3; ++ch) {
            float sc = exp2f((float)l);
            float x = dir[ch] * sc;
            float g = __bfloat162float(gin[3 + half * 3 * d.deg_view + l * 3 + ch]);
            u[ch] += g * cosf(half ? x + 1.57079637050628662109375f : x) * sc;
          }
      c = 3 + 6 * d.deg_view;
    }
    float a_n[3] = {0.f, 0.f, 0.f};     // dL/d n (normals_to_use)
    if (d.use_n_dot_v) {
      float gq = __bfloat162float(gin[c]);
#pragma unroll
      for (int i = 0; i < 3; ++i) a_n[i] += gq * v[i];
    }
    if (d.use_reflections) {
      // dir = v - 2 (n.v) n  ->  dL/dn = -2 (u.n) v - 2 (n.v) u
      float un = u[0] * n[0] + u[1] * n[1] + u[2] * n[2];
#pragma unroll
      for (int i = 0; i < 3; ++i) a_n[i] += -2.f * un * v[i] - 2.f * ndv * u[i];
    }
    float a_p[3] = {0.f, 0.f, 0.f}, a_d[3] = {0.f, 0.f, 0.f};
    if (d.use_pred_normals) { a_p[0] = a_n[0]; a_p[1] = a_n[1]; a_p[2] = a_n[2]; }
    else { a_d[0] = a_n[0]; a_d[1] = a_n[1]; a_d[2] = a_n[2]; }
    // ---- losses on the normals (weights are differentiated through extra_dw)
    const float w = weights[m];
    if (L.orient_mult > 0.f) {
      const float* no = L.orient_on_pred ? np_ : nd;
      float* ao = L.orient_on_pred ? a_p : a_d;
      float p = -(no[0] * v[0] + no[1] * v[1] + no[2] * v[2]);
      float pm = fminf(0.f, p);
      st_or += L.orient_mult * w * pm * pm;
      if (p < 0.f) {
#pragma unroll
        for (int i = 0; i < 3; ++i) ao[i] += L.orient_mult * w * 2.f * p * (-v[i]);
      }
    }
    if (L.prednorm_mult > 0.f) {
      float dot = nd[0] * np_[0] + nd[1] * np_[1] + nd[2] * np_[2];
      st_pn += L.prednorm_mult * w * (1.f - dot);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        a_p[i] += -L.prednorm_mult * w * nd[i];
        a_d[i] += -L.prednorm_mult * w * np_[i];
      }
    }
    float hg[11];                        // head gradients, in the column order of Wcat (models.py layout)
#pragma unroll
    for (int i = 0; i < 11; ++i) hg[i] = 0.f;
    hg[0] = d_raw_density ? d_raw_density[m] : 0.f;
    if (d.use_pred_normals) {
      float dg[3];
      neg_normalize_bwd(np_, s_p, cl_p, a_p, dg);
      d_grad_pred[m * 3] = dg[0]; d_grad_pred[m * 3 + 1] = dg[1]; d_grad_pred[m * 3 + 2] = dg[2];
      hg[1] = dg[0]; hg[2] = dg[1]; hg[3] = dg[2];
    }
    if (d_raw_diffuse) { hg[4] = d_raw_diffuse[m * 3]; hg[5] = d_raw_diffuse[m * 3 + 1]; hg[6] = d_raw_diffuse[m * 3 + 2]; }
    if (d_raw_tint) { hg[7] = d_raw_tint[m * 3]; hg[8] = d_raw_tint[m * 3 + 1]; hg[9] = d_raw_tint[m * 3 + 2]; }
    if (d.use_density_normals) {
      float dg[3];
      neg_normalize_bwd(nd, s_d, cl_d, a_d, dg);
      d_raw_grad_density[m] = dg[0]; d_raw_grad_density[d.M + m] = dg[1]; d_raw_grad_density[2 * d.M + m] = dg[2];
    }
    if (d.use_roughness) { hg[10] = dkappa * sigmoid_f(rin); d_raw_rough[m] = hg[10]; }
    // the consumed direction-encoding gradient columns are re-used for the head gradients: together
    // with the bottleneck gradient in columns [0, col0) they form the A operand of one dgrad GEMM
    __nv_bfloat16* hs = d_slab + m * (int64_t)ld_dslab + d.col0;
#pragma unroll
    for (int i = 0; i < 11; ++i) hs[i] = __float2bfloat16(hg[i]);
    for (int i = 11; d.col0 + i < d.col_end; ++i) hs[i] = __float2bfloat16(0.f);
  }
  st_or = warp_sum(st_or);
  st_pn = warp_sum(st_pn);
  if ((threadIdx.x & 31) == 0) {
    if (st_or != 0.f) atomicAdd(&stats[4], st_or);
    if (st_pn != 0.f) atomicAdd(&stats[5], st_pn);
  }
}

// out[r, n] (bf16) = mask(r mod mod, n) ? rowv[r] * colv[n] : 0     (start of the tangent backward chain)
__global__ void outer_mask_kernel(int64_t R, int N, int64_t mod, const float* __restrict__ rowv,
                                  const float* __restrict__ colv, const uint32_t* __restrict__ maskbits,
                                  int64_t ldmb, __nv_bfloat16* __restrict__ out, int64_t ldo) {
  const int64_t total = R * (N / 8);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / (N / 8);
    int c8 = (int)(i - r * (N / 8)) * 8;
    float rv = rowv[r];
    uint32_t bits = maskbits ? maskbits[(mod ? r % mod : r) * ldmb + (c8 >> 5)] >> (c8 & 31) : 0xffu;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = ((bits >> e) & 1u) ? rv * colv[c8 + e] : 0.f;
    uint4 o;
    o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + r * ldo + c8) = o;
  }
}

}  // namespace mnrf

extern "C" int mnrf_refdir_fwd(const mnrf_refdir_desc* d, const float* ide_mat, const int32_t* ide_ml,
                               const float* grad_pred, const float* raw_rough, const float* raw_grad_density,
                               const float* viewdirs, float* normals_pred, float* normals, float* roughness,
                               mnrf_bf16* slab, float orient_mult, float prednorm_mult, int32_t orient_on_pred,
                               float* extra_dw, mnrf_stream stream) {
  using namespace mnrf;
  if (d && d->M == 0) return 0;
  MNRF_CHECK(d && viewdirs && slab, "mnrf_refdir_fwd: null pointer");
  MNRF_CHECK(!d->use_ide || (ide_mat && ide_ml && d->ide_n <= kIdeMax && d->deg_view >= 1 && d->deg_view <= 5),
             "Only deg_view of at most 5 is numerically stable.");
  MNRF_CHECK(!d->use_ide || d->use_roughness, "mnrf_refdir_fwd: the IDE needs a roughness (kappa_inv)");
  MNRF_CHECK(d->use_pred_normals || d->use_density_normals || !(d->use_reflections || d->use_n_dot_v),
             "Normals must be computed for reflection directions.");
  if (d->M == 0) return 0;
  RefDesc r{d->M, d->num_samples, d->use_pred_normals, d->use_density_normals, d->use_reflections, d->use_ide,
            d->use_n_dot_v, d->use_roughness, d->deg_view, d->roughness_bias, d->ld, d->col0, d->col_end};
  int blocks = (int)std::min<int64_t>((d->M + 127) / 128, (int64_t)mnrf_num_sms() * 16);
  refdir_fwd_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(
      r, ide_mat, ide_ml, d->ide_n, grad_pred, raw_rough, raw_grad_density, viewdirs, normals_pred, normals,
      roughness, reinterpret_cast<__nv_bfloat16*>(slab), RefLoss{orient_mult, prednorm_mult, orient_on_pred},
      extra_dw);
  MNRF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mnrf_refdir_bwd(const mnrf_refdir_desc* d, const float* ide_mat, const int32_t* ide_ml,
                               const float* grad_pred, const float* raw_rough, const float* raw_grad_density,
                               const float* viewdirs, const float* weights, mnrf_bf16* d_slab,
                               int32_t ld_dslab, float orient_mult, float prednorm_mult, int32_t orient_on_pred,
                               const float* d_raw_density, const float* d_raw_diffuse, const float* d_raw_tint,
                               float* d_grad_pred, float* d_raw_rough, float* d_raw_grad_density,
                               float* stats, mnrf_stream stream) {
  using namespace mnrf;
  if (d && d->M == 0) return 0;
  MNRF_CHECK(d && viewdirs && weights && d_slab && stats, "mnrf_refdir_bwd: null pointer");
  MNRF_CHECK(d->col_end - d->col0 >= 11, "mnrf_refdir_bwd: the slab must hold the 11 head gradients");
  MNRF_CHECK(!d->use_ide || (ide_mat && ide_ml && d->ide_n <= kIdeMax), "mnrf_refdir_bwd: bad IDE tables");
  if (d->M == 0) return 0;
  RefDesc r{d->M, d->num_samples, d->use_pred_normals, d->use_density_normals, d->use_reflections, d->use_ide,
            d->use_n_dot_v, d->use_roughness, d->deg_view, d->roughness_bias, d->ld, d->col0, d->col_end};
  RefLoss L{orient_mult, prednorm_mult, orient_on_pred};
  int blocks = (int)std::min<int64_t>((d->M + 127) / 128, (int64_t)mnrf_num_sms() * 16);
  refdir_bwd_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(
      r, L, ide_mat, ide_ml, d->ide_n, grad_pred, raw_rough, raw_grad_density, viewdirs, weights,
      reinterpret_cast<__nv_bfloat16*>(d_slab), ld_dslab, d_raw_density, d_raw_diffuse, d_raw_tint, d_grad_pred,
      d_raw_rough, d_raw_grad_density, stats);
  MNRF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mnrf_outer_mask(int64_t rows, int32_t n, int64_t mask_mod, const float* rowv, const float* colv,
                               const uint32_t* maskbits, int64_t ldmaskbits, mnrf_bf16* out, int64_t ldo,
                               mnrf_stream stream) {
  using namespace mnrf;
  if (rows == 0) return 0;
  MNRF_CHECK(rowv && colv && out, "mnrf_outer_mask: null pointer");
  MNRF_CHECK(n % 32 == 0 && ldo % 8 == 0, "mnrf_outer_mask: N %% 32 == 0 and ld %% 8 == 0 required");
  if (rows == 0) return 0;
  int64_t total = rows * (n / 8);
  int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)mnrf_num_sms() * 16);
  outer_mask_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(rows, n, mask_mod, rowv, colv, maskbits, ldmaskbits,
                                                            reinterpret_cast<__nv_bfloat16*>(out), ldo);
  MNRF_LAUNCH_CHECK();
  return 0;
}
