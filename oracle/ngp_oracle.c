/* TEST INFRASTRUCTURE ONLY — never imported, linked or executed by the product path.
 *
 * CPU restatement (plain C, one serial loop per ray exactly like the reference's one thread per
 * ray) of the `vren` half of the instant-ngp-pp hot path, SURVEY.md §8 rows a1-a9.  Each function
 * cites the reference lines it follows.  Used by tests/ as the checker, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
 *
 * PINNING: tests/golden/*.npz hold outputs of the reference's own CUDA kernels (oracle/_ref/
 * vren_ref.so, built in place from /root/reference/models/csrc) on seeded inputs, minted on a
 * B200 by tests/golden/make_golden.py; tests/test_oracle_golden.py checks this file against them
 * (bit-exact for morton / packbits / AABB / marching, tolerance for the __expf compositors).
 *
 * Floating-point discipline: built with -ffp-contract=off; an FMA appears only where the
 * reference's sm_100 SASS has an FFMA whose fusion changes the rounded result
 * (x=o+t*d, x*inv+1, (..)*mip_bound-x, t1+=dt*noise — see DESIGN.md), spelt fmaf().
 *
 * Sample order: the reference's rays_a row order / start_idx come from two independent atomics
 * (raymarching.cu:237-241) and are nondeterministic; the canonical order used here (and by the
 * CUDA path) is ray-index order with start_idx = exclusive scan of N_samples.  Tests compare with
 * the reference after sorting its rays_a by column 0 and gathering segments.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SQRT3 1.73205080757f /* raymarching.cu:4 */
#define API __attribute__((visibility("default")))

static inline float clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); } /* helper_math.h:280 */

/* ------------------------------------------------------------------ intersection.cu:5-56 */
static void aabb1(const float* o, const float* d, const float* c, const float* h, float* t1o, float* t2o) {
  float tmin[3], tmax[3];
  for (int k = 0; k < 3; k++) {
    const float inv = 1.0f / d[k];
    tmin[k] = ((c[k] - h[k]) - o[k]) * inv;
    tmax[k] = ((c[k] + h[k]) - o[k]) * inv;
  }
  const float t1 = fmaxf(fmaxf(fminf(tmin[0], tmax[0]), fminf(tmin[1], tmax[1])), fminf(tmin[2], tmax[2]));
  const float t2 = fminf(fminf(fmaxf(tmin[0], tmax[0]), fmaxf(tmin[1], tmax[1])), fmaxf(tmin[2], tmax[2]));
  if (t1 > t2) { *t1o = -1.0f; *t2o = -1.0f; } else { *t1o = t1; *t2o = t2; }
}

/* intersection.cu:59-100 incl. the sort-by-t1 of all max_hits slots (-1 fillers sort first). */
API void ref_ray_aabb_intersect(const float* rays_o, const float* rays_d, const float* centers,
                                const float* half_sizes, int64_t n_rays, int64_t n_vox, int max_hits,
                                int32_t* hit_cnt, float* hits_t, int64_t* hits_idx) {
  for (int64_t r = 0; r < n_rays; r++) {
    float* row_t = hits_t + r * max_hits * 2;
    int64_t* row_i = hits_idx + r * max_hits;
    for (int k = 0; k < max_hits; k++) { row_t[2 * k] = -1.f; row_t[2 * k + 1] = -1.f; row_i[k] = -1; }
    int cnt = 0;
    for (int64_t v = 0; v < n_vox; v++) {
      float t1, t2;
      aabb1(rays_o + 3 * r, rays_d + 3 * r, centers + 3 * v, half_sizes + 3 * v, &t1, &t2);
      if (t2 > 0) {
        if (cnt < max_hits) { row_t[2 * cnt] = fmaxf(t1, 0.0f); row_t[2 * cnt + 1] = t2; row_i[cnt] = v; }
        cnt++;
      }
    }
    hit_cnt[r] = cnt;
    for (int i = 1; i < max_hits; i++) { /* stable insertion sort on t1 */
      const float k1 = row_t[2 * i], k2 = row_t[2 * i + 1];
      const int64_t ki = row_i[i];
      int j = i - 1;
      while (j >= 0 && row_t[2 * j] > k1) {
        row_t[2 * j + 2] = row_t[2 * j]; row_t[2 * j + 3] = row_t[2 * j + 1]; row_i[j + 1] = row_i[j]; j--;
      }
      row_t[2 * j + 2] = k1; row_t[2 * j + 3] = k2; row_i[j + 1] = ki;
    }
  }
}

/* ------------------------------------------------------------------ raymarching.cu:35-60 */
static inline uint32_t expand_bits(uint32_t v) {
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}
static inline uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) {
  return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
static inline uint32_t morton3D_invert(uint32_t x) {
  x = x & 0x49249249u;
  x = (x | (x >> 2)) & 0xc30c30c3u;
  x = (x | (x >> 4)) & 0x0f00f00fu;
  x = (x | (x >> 8)) & 0xff0000ffu;
  x = (x | (x >> 16)) & 0x0000ffffu;
  return x;
}
/* raymarching.cu:62-70 */
API void ref_morton3D(const int32_t* coords, int64_t n, int32_t* indices) {
  for (int64_t i = 0; i < n; i++)
    indices[i] = (int32_t)morton3D((uint32_t)coords[3 * i], (uint32_t)coords[3 * i + 1], (uint32_t)coords[3 * i + 2]);
}
/* raymarching.cu:90-101 */
API void ref_morton3D_invert(const int32_t* indices, int64_t n, int32_t* coords) {
  for (int64_t i = 0; i < n; i++) {
    const int32_t ind = indices[i];
    coords[3 * i + 0] = (int32_t)morton3D_invert((uint32_t)(ind >> 0));
    coords[3 * i + 1] = (int32_t)morton3D_invert((uint32_t)(ind >> 1));
    coords[3 * i + 2] = (int32_t)morton3D_invert((uint32_t)(ind >> 2));
  }
}
/* raymarching.cu:122-141 (float grid) */
API void ref_packbits(const float* grid, int64_t n_bytes, float thr, uint8_t* bitfield) {
  for (int64_t n = 0; n < n_bytes; n++) {
    uint8_t bits = 0;
    for (int i = 0; i < 8; i++) bits |= (grid[8 * n + i] > thr) ? (uint8_t)(1u << i) : 0;
    bitfield[n] = bits;
  }
}

/* ------------------------------------------------------------------ raymarching.cu:11-32 */
static inline float calc_dt(float t, float esf, int max_samples, int grid_size, float scale) {
  return clampf(t * esf, SQRT3 / max_samples, SQRT3 * 2 * scale / grid_size);
}
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int mip_from_pos(float x, float y, float z, int cascades) {
  const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
  int e; frexpf(mx, &e);
  return imin(cascades - 1, imax(0, e + 1));
}
static inline int mip_from_dt(float dt, int grid_size, int cascades) {
  int e; frexpf(dt * grid_size, &e);
  return imin(cascades - 1, imax(0, e));
}

typedef struct {
  const uint8_t* bitfield; int cascades, grid_size, max_samples; float scale, dt_scale, esf;
} march_cfg;

/* One step of raymarching.cu:205-232; returns 1 if the cell is occupied (sample at x,y,z with
 * step *dt), else advances *t past the empty cell. */
static int march_step(const march_cfg* c, const float* o, const float* d, const float* dinv, float* t,
                      float* xyz, float* dt) {
  const int G = c->grid_size;
  const float x = fmaf(*t, d[0], o[0]), y = fmaf(*t, d[1], o[1]), z = fmaf(*t, d[2], o[2]);
  *dt = calc_dt(*t, c->esf, c->max_samples, G, c->dt_scale);
  const int mip = imax(mip_from_pos(x, y, z, c->cascades), mip_from_dt(*dt, G, c->cascades));
  const float mip_bound = fminf(scalbnf(1.0f, mip - 1), c->scale);
  const float mip_bound_inv = 1 / mip_bound;
  const int nx = (int)clampf(0.5f * fmaf(x, mip_bound_inv, 1.0f) * G, 0.0f, G - 1.0f);
  const int ny = (int)clampf(0.5f * fmaf(y, mip_bound_inv, 1.0f) * G, 0.0f, G - 1.0f);
  const int nz = (int)clampf(0.5f * fmaf(z, mip_bound_inv, 1.0f) * G, 0.0f, G - 1.0f);
  const uint32_t idx = (uint32_t)mip * (uint32_t)(G * G * G) + morton3D(nx, ny, nz);
  const int occ = c->bitfield[idx / 8] & (1 << (idx % 8));
  xyz[0] = x; xyz[1] = y; xyz[2] = z;
  if (occ) return 1;
  const float ginv = 1.0f / G;
  const float tx = fmaf(((nx + 0.5f + 0.5f * copysignf(1.0f, d[0])) * ginv) * 2 - 1, mip_bound, -x) * dinv[0];
  const float ty = fmaf(((ny + 0.5f + 0.5f * copysignf(1.0f, d[1])) * ginv) * 2 - 1, mip_bound, -y) * dinv[1];
  const float tz = fmaf(((nz + 0.5f + 0.5f * copysignf(1.0f, d[2])) * ginv) * 2 - 1, mip_bound, -z) * dinv[2];
  const float t_target = *t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
  do { *t += calc_dt(*t, c->esf, c->max_samples, G, c->dt_scale); } while (*t < t_target);
  return 0;
}

/* raymarching.cu:166-280 in canonical (ray-index) order.  Pass `xyzs == NULL` to count only.
 * Returns total samples. rays_a (R,3) i64, xyzs/dirs (S,3), deltas/ts (S). */
API int64_t ref_raymarching_train(const float* rays_o, const float* rays_d, const float* hits_t,
                                  const uint8_t* bitfield, int cascades, float scale, float esf, const float* noise,
                                  int grid_size, int max_samples, int64_t n_rays, int64_t* rays_a, float* xyzs,
                                  float* dirs, float* deltas, float* ts) {
  const march_cfg c = {bitfield, cascades, grid_size, max_samples, scale, scale, esf};
  int64_t total = 0;
  for (int64_t r = 0; r < n_rays; r++) {
    const float* o = rays_o + 3 * r; const float* d = rays_d + 3 * r;
    const float dinv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
    float t1 = hits_t[2 * r]; const float t2 = hits_t[2 * r + 1];
    if (t1 >= 0) { const float dt = calc_dt(t1, esf, max_samples, grid_size, scale); t1 = fmaf(dt, noise[r], t1); }
    float t = t1, xyz[3], dt; int N = 0;
    while (0 <= t && t < t2 && N < max_samples) {
      if (march_step(&c, o, d, dinv, &t, xyz, &dt)) { t += dt; N++; }
    }
    rays_a[3 * r] = r; rays_a[3 * r + 1] = total; rays_a[3 * r + 2] = N;
    if (xyzs) {
      t = t1; int s = 0;
      while (t < t2 && s < N) {
        const float tt = t;
        if (march_step(&c, o, d, dinv, &t, xyz, &dt)) {
          const int64_t q = total + s;
          memcpy(xyzs + 3 * q, xyz, 12); memcpy(dirs + 3 * q, d, 12);
          ts[q] = tt; deltas[q] = dt; t += dt; s++;
        }
      }
    }
    total += N;
  }
  return total;
}

/* raymarching.cu:335-404, incl. calc_dt(..., cascades) (:370,399) and the in-place hits_t update. */
API void ref_raymarching_test(const float* rays_o, const float* rays_d, float* hits_t, const int64_t* alive,
                              const uint8_t* bitfield, int cascades, float scale, float esf, int grid_size,
                              int max_samples, int n_samples, int64_t n_alive, float* xyzs, float* dirs,
                              float* deltas, float* ts, int32_t* n_eff) {
  const march_cfg c = {bitfield, cascades, grid_size, max_samples, scale, (float)cascades, esf};
  memset(xyzs, 0, sizeof(float) * 3 * n_alive * n_samples); memset(dirs, 0, sizeof(float) * 3 * n_alive * n_samples);
  memset(deltas, 0, sizeof(float) * n_alive * n_samples); memset(ts, 0, sizeof(float) * n_alive * n_samples);
  for (int64_t n = 0; n < n_alive; n++) {
    const int64_t r = alive[n];
    const float* o = rays_o + 3 * r; const float* d = rays_d + 3 * r;
    const float dinv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
    float t = hits_t[2 * r]; const float t2 = hits_t[2 * r + 1];
    float xyz[3], dt; int s = 0;
    while (t < t2 && s < n_samples) {
      const float tt = t;
      if (march_step(&c, o, d, dinv, &t, xyz, &dt)) {
        const int64_t q = n * n_samples + s;
        memcpy(xyzs + 3 * q, xyz, 12); memcpy(dirs + 3 * q, d, 12);
        ts[q] = tt; deltas[q] = dt; t += dt; hits_t[2 * r] = t; s++;
      }
    }
    n_eff[n] = s;
  }
}

/* ------------------------------------------------------------------ volumerendering.cu
 * __expf(x) is ex2.approx(x*log2e) on the GPU; the oracle uses expf — compositor parity is a
 * tolerance check (tests state it). */
static inline float alpha_of(float sigma, float delta) { return 1.0f - expf(-sigma * delta); }

/* volumerendering.cu:5-34 */
API void ref_composite_alpha_fw(const float* sigmas, const float* deltas, const int64_t* rays_a, float T_thr,
                                int64_t n_samples, int64_t n_rays, float* alphas, float* ws) {
  memset(alphas, 0, sizeof(float) * n_samples); memset(ws, 0, sizeof(float) * n_samples);
  for (int64_t n = 0; n < n_rays; n++) {
    const int64_t start = rays_a[3 * n + 1]; const int N = (int)rays_a[3 * n + 2];
    int samples = 0; float T = 1.0f;
    while (samples < N) {
      const int64_t s = start + samples;
      const float a = alpha_of(sigmas[s], deltas[s]); const float w = a * T;
      alphas[s] = a; ws[s] = w; T *= 1.0f - a;
      if (T <= T_thr) break;
      samples++;
    }
  }
}

/* volumerendering.cu:65-115 */
API void ref_composite_train_fw(const float* sigmas, const float* rgbs, const float* normals_pred, const float* sems,
                                const float* deltas, const float* ts, const int64_t* rays_a, float T_thr, int classes,
                                int64_t n_samples, int64_t n_rays, int64_t* total_samples, float* opacity,
                                float* depth, float* rgb, float* normal_pred, float* sem, float* ws) {
  memset(total_samples, 0, sizeof(int64_t) * n_rays); memset(opacity, 0, sizeof(float) * n_rays);
  memset(depth, 0, sizeof(float) * n_rays); memset(rgb, 0, sizeof(float) * 3 * n_rays);
  memset(normal_pred, 0, sizeof(float) * 3 * n_rays); memset(sem, 0, sizeof(float) * classes * n_rays);
  memset(ws, 0, sizeof(float) * n_samples);
  for (int64_t n = 0; n < n_rays; n++) {
    const int64_t ray = rays_a[3 * n], start = rays_a[3 * n + 1]; const int N = (int)rays_a[3 * n + 2];
    int samples = 0; float T = 1.0f;
    while (samples < N) {
      const int64_t s = start + samples;
      const float a = alpha_of(sigmas[s], deltas[s]); const float w = a * T;
      for (int k = 0; k < 3; k++) rgb[3 * ray + k] = fmaf(w, rgbs[3 * s + k], rgb[3 * ray + k]);
      for (int k = 0; k < 3; k++) normal_pred[3 * ray + k] = fmaf(w, normals_pred[3 * s + k], normal_pred[3 * ray + k]);
      depth[ray] = fmaf(w, ts[s], depth[ray]);
      for (int i = 0; i < classes; i++) sem[ray * classes + i] = fmaf(w, sems[s * classes + i], sem[ray * classes + i]);
      opacity[ray] += w; ws[s] = w; T *= 1.0f - a;
      if (T <= T_thr) break;
      samples++;
    }
    total_samples[ray] = samples;
  }
}

/* volumerendering.cu:167-311 (incl. host-side dL_dws*ws :277 and the sequential inclusive scan :206-210) */
API void ref_composite_train_bw(const float* dL_dopacity, const float* dL_ddepth, const float* dL_drgb,
                                const float* dL_dnormal_pred, const float* dL_dsem, const float* dL_dws,
                                const float* sigmas, const float* rgbs, const float* ws, const float* deltas,
                                const float* ts, const int64_t* rays_a, const float* opacity, const float* depth,
                                const float* rgb, float T_thr, int classes, int64_t n_samples, int64_t n_rays,
                                float* dL_dsigmas, float* dL_drgbs, float* dL_dnormals_pred, float* dL_dsems) {
  memset(dL_dsigmas, 0, sizeof(float) * n_samples); memset(dL_drgbs, 0, sizeof(float) * 3 * n_samples);
  memset(dL_dnormals_pred, 0, sizeof(float) * 3 * n_samples); memset(dL_dsems, 0, sizeof(float) * classes * n_samples);
  float* scan = (float*)malloc(sizeof(float) * (n_samples > 0 ? n_samples : 1));
  for (int64_t i = 0; i < n_samples; i++) scan[i] = dL_dws[i] * ws[i];
  for (int64_t n = 0; n < n_rays; n++) {
    const int64_t ray = rays_a[3 * n], start = rays_a[3 * n + 1]; const int N = (int)rays_a[3 * n + 2];
    if (N <= 0) continue;
    for (int k = 1; k < N; k++) scan[start + k] += scan[start + k - 1];
    const float sum = scan[start + N - 1];
    const float R = rgb[3 * ray], G = rgb[3 * ray + 1], B = rgb[3 * ray + 2], O = opacity[ray], D = depth[ray];
    float T = 1.0f, r = 0, g = 0, b = 0, d = 0; int samples = 0;
    while (samples < N) {
      const int64_t s = start + samples;
      const float a = alpha_of(sigmas[s], deltas[s]); const float w = a * T;
      r = fmaf(w, rgbs[3 * s], r); g = fmaf(w, rgbs[3 * s + 1], g); b = fmaf(w, rgbs[3 * s + 2], b); d = fmaf(w, ts[s], d);
      T *= 1.0f - a;
      for (int k = 0; k < 3; k++) dL_drgbs[3 * s + k] = dL_drgb[3 * ray + k] * w;
      for (int k = 0; k < 3; k++) dL_dnormals_pred[3 * s + k] = dL_dnormal_pred[3 * ray + k] * w;
      for (int i = 0; i < classes; i++) dL_dsems[s * classes + i] = dL_dsem[ray * classes + i] * w;
      dL_dsigmas[s] = deltas[s] * (dL_drgb[3 * ray] * (rgbs[3 * s] * T - (R - r)) +
                                   dL_drgb[3 * ray + 1] * (rgbs[3 * s + 1] * T - (G - g)) +
                                   dL_drgb[3 * ray + 2] * (rgbs[3 * s + 2] * T - (B - b)) +
                                   dL_dopacity[ray] * (1 - O) + dL_ddepth[ray] * (ts[s] * T - (D - d)) +
                                   T * dL_dws[s] - (sum - scan[s]));
      if (T <= T_thr) break;
      samples++;
    }
  }
  free(scan);
}

/* volumerendering.cu:314-374 */
API void ref_composite_test_fw(const float* sigmas, const float* rgbs, const float* normals, const float* normals_raw,
                               const float* sems, const float* deltas, const float* ts, int64_t* alive, float T_thr,
                               int classes, const int32_t* n_eff, int n_samples, int64_t n_alive, float* opacity,
                               float* depth, float* rgb, float* normal, float* normal_raw, float* sem) {
  for (int64_t n = 0; n < n_alive; n++) {
    if (n_eff[n] == 0) { alive[n] = -1; continue; }
    const int64_t r = alive[n];
    int s = 0; float T = 1 - opacity[r];
    while (s < n_eff[n]) {
      const int64_t q = n * n_samples + s;
      const float a = alpha_of(sigmas[q], deltas[q]); const float w = a * T;
      for (int k = 0; k < 3; k++) rgb[3 * r + k] = fmaf(w, rgbs[3 * q + k], rgb[3 * r + k]);
      depth[r] = fmaf(w, ts[q], depth[r]); opacity[r] += w;
      for (int k = 0; k < 3; k++) normal[3 * r + k] = fmaf(w, normals[3 * q + k], normal[3 * r + k]);
      for (int k = 0; k < 3; k++) normal_raw[3 * r + k] = fmaf(w, normals_raw[3 * q + k], normal_raw[3 * r + k]);
      for (int i = 0; i < classes; i++) sem[r * classes + i] = fmaf(w, sems[q * classes + i], sem[r * classes + i]);
      T *= 1.0f - a;
      if (T <= T_thr) { alive[n] = -1; break; }
      s++;
    }
  }
}

/* ------------------------------------------------------------------ ref_loss.cu:4-38 */
API void ref_composite_refloss_fw(const float* sigmas, const float* ndiff, const float* nori, const float* deltas,
                                  const int64_t* rays_a, float T_thr, int64_t n_samples, int64_t n_rays,
                                  float* loss_o, float* loss_p) {
  (void)n_samples;
  memset(loss_o, 0, sizeof(float) * n_rays); memset(loss_p, 0, sizeof(float) * 3 * n_rays);
  for (int64_t n = 0; n < n_rays; n++) {
    const int64_t ray = rays_a[3 * n], start = rays_a[3 * n + 1]; const int N = (int)rays_a[3 * n + 2];
    int samples = 0; float T = 1.0f;
    while (samples < N) {
      const int64_t s = start + samples;
      const float a = alpha_of(sigmas[s], deltas[s]); const float w = a * T;
      for (int k = 0; k < 3; k++) loss_p[3 * ray + k] = fmaf(w, ndiff[3 * s + k], loss_p[3 * ray + k]);
      loss_o[ray] = fmaf(w, nori[s], loss_o[ray]);
      T *= 1.0f - a;
      if (T <= T_thr) break;
      samples++;
    }
  }
}

/* ref_loss.cu:76-130 */
API void ref_composite_refloss_bw(const float* dL_dloss_o, const float* dL_dloss_p, const float* sigmas,
                                  const float* ndiff, const float* nori, const float* deltas, const int64_t* rays_a,
                                  const float* loss_o, const float* loss_p, float T_thr, int64_t n_samples,
                                  int64_t n_rays, float* dL_dsigmas, float* dL_dndiff, float* dL_dnori) {
  memset(dL_dsigmas, 0, sizeof(float) * n_samples); memset(dL_dndiff, 0, sizeof(float) * 3 * n_samples);
  memset(dL_dnori, 0, sizeof(float) * n_samples);
  for (int64_t n = 0; n < n_rays; n++) {
    const int64_t ray = rays_a[3 * n], start = rays_a[3 * n + 1]; const int N = (int)rays_a[3 * n + 2];
    const float X = loss_p[3 * ray], Y = loss_p[3 * ray + 1], Z = loss_p[3 * ray + 2], O = loss_o[ray];
    float T = 1.0f, x = 0, y = 0, z = 0, o = 0; int samples = 0;
    while (samples < N) {
      const int64_t s = start + samples;
      const float a = alpha_of(sigmas[s], deltas[s]); const float w = a * T;
      x = fmaf(w, ndiff[3 * s], x); y = fmaf(w, ndiff[3 * s + 1], y); z = fmaf(w, ndiff[3 * s + 2], z);
      o = fmaf(w, nori[s], o);
      T *= 1.0f - a;
      for (int k = 0; k < 3; k++) dL_dndiff[3 * s + k] = dL_dloss_p[3 * ray + k] * w;
      dL_dnori[s] = dL_dloss_o[ray] * w;
      dL_dsigmas[s] = deltas[s] * (dL_dloss_p[3 * ray] * (ndiff[3 * s] * T - (X - x)) +
                                   dL_dloss_p[3 * ray + 1] * (ndiff[3 * s + 1] * T - (Y - y)) +
                                   dL_dloss_p[3 * ray + 2] * (ndiff[3 * s + 2] * T - (Z - z)) +
                                   dL_dloss_o[ray] * (nori[s] * T - (O - o)));
      if (T <= T_thr) break;
      samples++;
    }
  }
}

/* ------------------------------------------------------------------ losses.cu:7-107 */
API void ref_distortion_loss_fw(const float* ws, const float* deltas, const float* ts, const int64_t* rays_a,
                                int64_t n_samples, int64_t n_rays, float* loss, float* ws_incl, float* wts_incl) {
  memset(loss, 0, sizeof(float) * n_rays);
  memset(ws_incl, 0, sizeof(float) * n_samples); memset(wts_incl, 0, sizeof(float) * n_samples);
  for (int64_t n = 0; n < n_rays; n++) {
    const int64_t ray = rays_a[3 * n], start = rays_a[3 * n + 1]; const int N = (int)rays_a[3 * n + 2];
    float iw = 0, iwt = 0, acc = 0;
    for (int k = 0; k < N; k++) {
      const int64_t s = start + k;
      const float ew = iw, ewt = iwt;
      const float wt = ws[s] * ts[s];
      iw = (k == 0) ? ws[s] : iw + ws[s]; iwt = (k == 0) ? wt : iwt + wt;
      ws_incl[s] = iw; wts_incl[s] = iwt;
      const float l = 2 * (iwt * ew - iw * ewt) + 1.0f / 3 * ws[s] * ws[s] * deltas[s];
      acc += l;
    }
    loss[ray] = acc;
  }
}

/* losses.cu:110-140 */
API void ref_distortion_loss_bw(const float* dL_dloss, const float* ws_incl, const float* wts_incl, const float* ws,
                                const float* deltas, const float* ts, const int64_t* rays_a, int64_t n_samples,
                                int64_t n_rays, float* dL_dws) {
  memset(dL_dws, 0, sizeof(float) * n_samples);
  for (int64_t n = 0; n < n_rays; n++) {
    const int64_t ray = rays_a[3 * n], start = rays_a[3 * n + 1]; const int N = (int)rays_a[3 * n + 2];
    if (N <= 0) continue;
    const int64_t end = start + N - 1;
    const float ws_sum = ws_incl[end], wts_sum = wts_incl[end];
    for (int64_t s = start; s <= end; s++) {
      dL_dws[s] = dL_dloss[ray] * 2 *
                  ((s == start ? 0.0f : (ts[s] * ws_incl[s - 1] - wts_incl[s - 1])) +
                   (wts_sum - wts_incl[s] - ts[s] * (ws_sum - ws_incl[s])));
      dL_dws[s] += dL_dloss[ray] * (float)2 / 3 * ws[s] * deltas[s];
    }
  }
}
