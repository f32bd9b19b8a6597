/*
 * oracle/c/nrt_oracle.c -- C99 + OpenMP restatement of the reference's hot path.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for full-size parity and
 * the multi-threaded CPU baseline of bench.py.  Never linked into the product.
 *
 * Same arithmetic, op for op, as oracle/interp.py (which is pinned bit-exactly to the
 * reference's own source run on tools/tfshim): one fp32 rounding per operation -- compile
 * with -ffp-contract=off so the compiler cannot fuse a*b+c.
 *
 *   oracle_warp_f32      voxelmorph SpatialTransformer contract = identity grid + flow ->
 *                        interpn, reference neurite/tf/utils/utils.py:73-220
 *   oracle_interpn_f32   reference neurite/tf/utils/utils.py:73-220 (explicit loc)
 *   oracle_dice_sums_f32 reference neurite/tf/metrics.py:471-477 (the three reductions)
 *   oracle_cce_f32       reference neurite/tf/metrics.py:640-650 + Keras CCE formula
 *   oracle_lc3d_f32      reference neurite/tf/layers.py:1126-1197, 1098-1099
 *   oracle_mi_volumes_f32  reference neurite/tf/metrics.py:185-292 (channelwise -> maps) with
 *                        neurite/tf/utils/utils.py:1099-1172 soft_quantize fused in
 *   oracle_sepconv_axis_f32  one pass of neurite/tf/utils/utils.py:665-751 separable_conv
 *                        (tf.nn.convolution: zero-padded cross-correlation, third party)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXD 5

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static inline float clipf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* one output point; vol [S.., C] resident in full; loc[D] */
static void sample_point(const float* vol, const int* S, int D, int C, const float* loc, int method,
                         int has_fill, float fill, float* out) {
  if (method == 0) {
    int i0[MAXD], i1[MAXD];
    float wlo[MAXD], whi[MAXD];
    for (int d = 0; d < D; ++d) {
      const float mx = (float)(S[d] - 1);
      const float x = clipf(loc[d], 0.0f, mx);                 /* utils.py:142 */
      const float f0 = clipf(floorf(loc[d]), 0.0f, mx);        /* :139, :143   */
      const float f1 = clipf(f0 + 1.0f, 0.0f, mx);             /* :146         */
      i0[d] = (int)f0; i1[d] = (int)f1;                        /* :147         */
      wlo[d] = f1 - x;                                         /* :152 diff_loc1 -> corner bit 0 */
      whi[d] = 1.0f - wlo[d];                                  /* :153 diff_loc0 -> corner bit 1 */
    }
    for (int c = 0; c < C; ++c) out[c] = 0.0f;                 /* :160 */
    for (int corner = 0; corner < (1 << D); ++corner) {        /* itertools.product order, :159 */
      int idx = 0;
      float w = 0.0f;
      for (int d = 0; d < D; ++d) {
        const int bit = (corner >> (D - 1 - d)) & 1;
        idx = idx * S[d] + (bit ? i1[d] : i0[d]);              /* sub2ind2d :1068-1082 */
        const float wd = bit ? whi[d] : wlo[d];
        w = d == 0 ? wd : w * wd;                              /* prod_n :1085-1092 */
      }
      const float* v = vol + (size_t)idx * C;
      for (int c = 0; c < C; ++c) {
        const float prod = w * v[c];
        out[c] = out[c] + prod;                                /* :191 */
      }
    }
  } else {
    int idx = 0;
    for (int d = 0; d < D; ++d) {
      int r = (int)nearbyintf(loc[d]);                         /* tf.round half-to-even, then cast, :196 */
      r = r < 0 ? 0 : (r > S[d] - 1 ? S[d] - 1 : r);           /* :197 */
      idx = idx * S[d] + r;
    }
    const float* v = vol + (size_t)idx * C;
    for (int c = 0; c < C; ++c) out[c] = v[c];
  }
  if (has_fill) {                                              /* :206-213 */
    int oob = 0;
    for (int d = 0; d < D; ++d) oob |= (loc[d] < 0.0f) || (loc[d] > (float)(S[d] - 1));
    const float keep = oob ? 0.0f : 1.0f, o = oob ? 1.0f : 0.0f;
    for (int c = 0; c < C; ++c) {
      const float a = out[c] * keep, b = o * fill;
      out[c] = a + b;
    }
  }
}

/* dynamic chunks: on a shared host one descheduled thread would otherwise hold up the whole static loop (measured on
 * the GPU boxes: the same call took 5 ms or 100 ms); every point is independent, so the result does not depend on it */
void oracle_interpn_f32(const float* vol, const int* S, int D, int C, const float* loc, int64_t n,
                        int method, int has_fill, float fill, float* out) {
#pragma omp parallel for schedule(dynamic, 4096)
  for (int64_t i = 0; i < n; ++i)
    sample_point(vol, S, D, C, loc + i * D, method, has_fill, fill, out + i * C);
}

/* vol [B, S.., C], flow [B, S.., D] -> out [B, S.., C] */
void oracle_warp_f32(const float* vol, const float* flow, float* out, int B, const int* S, int D, int C,
                     int method, int has_fill, float fill) {
  int64_t nvox = 1;
  for (int d = 0; d < D; ++d) nvox *= S[d];
  for (int b = 0; b < B; ++b) {
    const float* vb = vol + (size_t)b * nvox * C;
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t v = 0; v < nvox; ++v) {
      int64_t rem = v;
      float loc[MAXD];
      int coord[MAXD];
      for (int d = D - 1; d >= 0; --d) { coord[d] = (int)(rem % S[d]); rem /= S[d]; }
      const float* f = flow + ((size_t)b * nvox + v) * D;
      for (int d = 0; d < D; ++d) loc[d] = (float)coord[d] + f[d];
      sample_point(vb, S, D, C, loc, method, has_fill, fill, out + ((size_t)b * nvox + v) * C);
    }
  }
}

/* Parallel copy with the static schedule of the compute loops: a buffer filled by this function is
 * first-touched by the threads that will later read it (NUMA placement of the CPU baseline's inputs). */
void oracle_first_touch_copy_f32(float* dst, const float* src, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) dst[i] = src[i];
}

/* sums[b][l][3] = {sum t*p, sum t*t, sum p*p}; double accumulation, one rounding.  Voxel-parallel: every thread
 * streams a contiguous voxel range once (all labels of a voxel are adjacent in channels-last memory) into
 * private double partials, combined in thread order -- deterministic for a fixed thread count. */
#define ORACLE_MAX_LABELS 1024
void oracle_dice_sums_f32(const float* t, const float* p, int B, int64_t V, int L, float* sums) {
  for (int b = 0; b < B; ++b) {
    const float* tb = t + (size_t)b * V * L;
    const float* pb = p + (size_t)b * V * L;
    double tot[ORACLE_MAX_LABELS * 3];
    const int Lc = L <= ORACLE_MAX_LABELS ? L : ORACLE_MAX_LABELS;
    for (int i = 0; i < Lc * 3; ++i) tot[i] = 0.0;
    if (L > ORACLE_MAX_LABELS) {                      /* not used by the tests / bench: plain serial fallback */
      for (int l = 0; l < L; ++l) {
        double tp = 0, tt = 0, pp = 0;
        for (int64_t v = 0; v < V; ++v) {
          const float a = tb[v * L + l], c = pb[v * L + l];
          tp += (double)(a * c); tt += (double)(a * a); pp += (double)(c * c);
        }
        sums[((size_t)b * L + l) * 3 + 0] = (float)tp;
        sums[((size_t)b * L + l) * 3 + 1] = (float)tt;
        sums[((size_t)b * L + l) * 3 + 2] = (float)pp;
      }
      continue;
    }
#pragma omp parallel
    {
      double acc[ORACLE_MAX_LABELS * 3];
      for (int i = 0; i < L * 3; ++i) acc[i] = 0.0;
#pragma omp for schedule(static) nowait
      for (int64_t v = 0; v < V; ++v) {
        const float* tv = tb + v * L;
        const float* pv = pb + v * L;
        for (int l = 0; l < L; ++l) {
          const float a = tv[l], c = pv[l];
          acc[l * 3 + 0] += (double)(a * c); acc[l * 3 + 1] += (double)(a * a); acc[l * 3 + 2] += (double)(c * c);
        }
      }
#pragma omp for ordered schedule(static, 1)
      for (int th = 0; th < omp_get_num_threads(); ++th) {
#pragma omp ordered
        for (int i = 0; i < L * 3; ++i) tot[i] += acc[i];
      }
    }
    for (int i = 0; i < L * 3; ++i) sums[(size_t)b * L * 3 + i] = (float)tot[i];
  }
}

/* returns sum over rows of -sum_c (w_c t_c) log(clip(p_c / sum p)) as double */
double oracle_cce_f32(const float* t, const float* p, const float* lw, int64_t n, int C) {
  double total = 0.0;
#pragma omp parallel for reduction(+ : total) schedule(static)
  for (int64_t r = 0; r < n; ++r) {
    const float* tr = t + r * C;
    const float* pr = p + r * C;
    double s = 0.0;
    for (int c = 0; c < C; ++c) s += pr[c];
    const float sf = (float)s;
    double l = 0.0;
    for (int c = 0; c < C; ++c) {
      float q = pr[c] / sf;
      q = clipf(q, 1e-7f, 1.0f - 1e-7f);
      const float tv = lw ? lw[c] * tr[c] : tr[c];
      l += (double)tv * log((double)q);
    }
    total += (double)(float)(-l);
  }
  return total;
}

/* x [B,I0,I1,I2,Cin], kernel [P,F,Cout], bias [P,Cout] or NULL -> out [B,P,Cout]; channels_last order */
void oracle_lc3d_f32(const float* x, const float* kernel, const float* bias, float* out, int B, const int* I,
                     int Cin, int Cout, const int* K, const int* St) {
  int O[3];
  for (int d = 0; d < 3; ++d) O[d] = (I[d] - K[d]) / St[d] + 1;
  const int64_t P = (int64_t)O[0] * O[1] * O[2];
  const int F = K[0] * K[1] * K[2] * Cin;
  const int64_t xb = (int64_t)I[0] * I[1] * I[2] * Cin;
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < P; ++p) {
    const int o2 = (int)(p % O[2]), o1 = (int)((p / O[2]) % O[1]), o0 = (int)(p / ((int64_t)O[2] * O[1]));
    const float* w = kernel + p * (int64_t)F * Cout;
    for (int b = 0; b < B; ++b) {
      double acc[256];
      for (int f = 0; f < Cout; ++f) acc[f] = 0.0;
      int j = 0;
      for (int i0 = 0; i0 < K[0]; ++i0)
        for (int i1 = 0; i1 < K[1]; ++i1)
          for (int i2 = 0; i2 < K[2]; ++i2)
            for (int c = 0; c < Cin; ++c, ++j) {
              const float xv = x[b * xb + ((((int64_t)(o0 * St[0] + i0) * I[1]) + (o1 * St[1] + i1)) * I[2] + (o2 * St[2] + i2)) * Cin + c];
              const float* wr = w + (int64_t)j * Cout;
              for (int f = 0; f < Cout; ++f) acc[f] += (double)xv * (double)wr[f];
            }
      for (int f = 0; f < Cout; ++f) {
        float r = (float)acc[f];
        if (bias) r = r + bias[p * Cout + f];
        out[((int64_t)b * P + p) * Cout + f] = r;
      }
    }
  }
}


/* ---- MutualInformation.channelwise (metrics.py:185-225) on [B, V, C] intensity tensors --------
 * bins = tf.linspace(min, max, nb) of the WHOLE tensor (utils.py:1151-1153) unless `centers` is
 * given for both; w[bin] = exp(-alpha (clip(x) - c)^2) (utils.py:1157-1171); joint and marginal
 * sums in double (TF's matmul / reduce order is unspecified); maps() arithmetic in fp32
 * (metrics.py:265-292).  mi[b*C + c]. */
static void linspace_f32(float mn, float mx, int nb, float* c) {
  if (nb == 1) { c[0] = mn; return; }
  const float delta = (mx - mn) / (float)(nb - 1);
  c[0] = mn;
  for (int i = 1; i < nb - 1; ++i) c[i] = mn + delta * (float)i;
  c[nb - 1] = mx;
}

void oracle_mi_channelwise_f32(const float* x, const float* y, int B, int64_t V, int C, int nb, float alpha,
                               float lo, float hi, const float* centers_x, const float* centers_y, float* mi) {
  float cx[64], cy[64];
  if (nb > 64) return;
  const int64_t n = (int64_t)B * V * C;
  if (centers_x && centers_y) {
    memcpy(cx, centers_x, sizeof(float) * nb);
    memcpy(cy, centers_y, sizeof(float) * nb);
  } else {
    float mnx = INFINITY, mxx = -INFINITY, mny = INFINITY, mxy = -INFINITY;
#pragma omp parallel for reduction(min : mnx, mny) reduction(max : mxx, mxy) schedule(static)
    for (int64_t e = 0; e < n; ++e) {
      mnx = fminf(mnx, x[e]); mxx = fmaxf(mxx, x[e]);
      mny = fminf(mny, y[e]); mxy = fmaxf(mxy, y[e]);
    }
    linspace_f32(mnx, mxx, nb, cx);
    linspace_f32(mny, mxy, nb, cy);
  }
  const float eps = 1e-7f;
  for (int item = 0; item < B * C; ++item) {
    const int b = item / C, c = item - b * C;
    double H[64 * 64], sx[64], sy[64];
    memset(H, 0, sizeof(H)); memset(sx, 0, sizeof(sx)); memset(sy, 0, sizeof(sy));
#pragma omp parallel
    {
      double h[64 * 64], px[64], py[64];
      memset(h, 0, sizeof(double) * nb * nb); memset(px, 0, sizeof(px)); memset(py, 0, sizeof(py));
#pragma omp for schedule(static) nowait
      for (int64_t v = 0; v < V; ++v) {
        const float xv = clipf(x[((int64_t)b * V + v) * C + c], lo, hi);
        const float yv = clipf(y[((int64_t)b * V + v) * C + c], lo, hi);
        float wx[64], wy[64];
        for (int i = 0; i < nb; ++i) {
          const float dx = xv - cx[i], dy = yv - cy[i];
          wx[i] = expf(-alpha * (dx * dx));
          wy[i] = expf(-alpha * (dy * dy));
          px[i] += wx[i]; py[i] += wy[i];
        }
        for (int i = 0; i < nb; ++i)
          for (int j = 0; j < nb; ++j) h[i * nb + j] += (double)wx[i] * (double)wy[j];
      }
#pragma omp critical
      {
        for (int i = 0; i < nb * nb; ++i) H[i] += h[i];
        for (int i = 0; i < nb; ++i) { sx[i] += px[i]; sy[i] += py[i]; }
      }
    }
    double th = 0, tx = 0, ty = 0;
    for (int i = 0; i < nb * nb; ++i) th += (double)(float)H[i];
    for (int i = 0; i < nb; ++i) { tx += (double)(float)sx[i]; ty += (double)(float)sy[i]; }
    const float N = (float)th + eps, Nx = (float)tx + eps, Ny = (float)ty + eps;
    double acc = 0;
    for (int i = 0; i < nb; ++i)
      for (int j = 0; j < nb; ++j) {
        const float pxy = (float)H[i * nb + j] / N;
        const float px_ = (float)sx[i] / Nx, py_ = (float)sy[j] / Ny;
        const float q = px_ * py_ + eps;
        acc += (double)(pxy * logf(pxy / q + eps));
      }
    mi[item] = (float)acc;
  }
}

/* one pass of separable_conv along the middle axis of [outer, L, inner] -> [outer, L_out, inner]:
 * out[o, l, i] = sum_j k[j] x[o, l*stride - pad_before + j*dil, i], zero outside; double accumulate. */
void oracle_sepconv_axis_f32(const float* x, float* out, int64_t outer, int64_t L, int64_t inner, const float* k, int K,
                             int stride, int dil, int pad_before, int64_t L_out) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t o = 0; o < outer; ++o)
    for (int64_t l = 0; l < L_out; ++l) {
      float* op = out + (o * L_out + l) * inner;
      for (int64_t i = 0; i < inner; ++i) {
        double acc = 0;
        for (int j = 0; j < K; ++j) {
          const int64_t s = l * stride - pad_before + (int64_t)j * dil;
          if (s >= 0 && s < L) acc += (double)k[j] * (double)x[(o * L + s) * inner + i];
        }
        op[i] = (float)acc;
      }
    }
}
