// extern "C" layer declared in include/b200_groth16.h: runtime plumbing, the O(1) host-side group operations of the
// prover tail, the device-resident proving key and the whole-proof entry points.
#include <algorithm>
#include <chrono>
#include <functional>
#include <condition_variable>
#include <future>
#include <memory>
#include <mutex>
#include <thread>
#include <cstdlib>
#include <cstring>
#include "../../include/b200_groth16.h"
#include "common.cuh"
#include "curve.cuh"
#include "devops.h"
#include "msm.h"
#include "msm_internal.h"
#include "ntt.h"

namespace b200 {
std::string &last_error() {
  static thread_local std::string e;
  return e;
}
int set_error(int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

std::atomic<unsigned long long> &launch_counter() {
  static std::atomic<unsigned long long> n{0};
  return n;
}

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- host group helpers, one instantiation per (curve, group) -------------------------------------------------
template <class G>
struct HostOps {
  typedef typename G::F F;
  typedef typename G::ScalarPrime FrP;
  static void add(const void *p, const void *q, void *out) {
    Proj<F> a, b, c;
    memcpy(&a, p, sizeof(a));
    memcpy(&b, q, sizeof(b));
    proj_add<G>(c, a, b);
    memcpy(out, &c, sizeof(c));
  }
  // scalar_mul(base, fr.as_bigint()) - curve_utils.tcc:13-34 via operator* (prover_reference_functions.cpp:166-168)
  static void scale(const void *fr, const void *p, void *out) {
    Fp<FrP> k;
    memcpy(&k, fr, sizeof(k));
    Fp<FrP>::from_mont(k, k);
    Proj<F> a, c;
    memcpy(&a, p, sizeof(a));
    // fixed 4-bit windows: 14 additions for the table 1P .. 15P, then 4 doublings + at most one addition per nibble
    // (753 doublings + ~190 additions instead of + ~376): the same group element, hence the same affine bytes
    Proj<F> tab[16];
    proj_set_zero(tab[0]);
    tab[1] = a;
    for (int i = 2; i < 16; i++) {
      if (i % 2 == 0) proj_dbl<G>(tab[i], tab[i / 2]);
      else proj_add<G>(tab[i], tab[i - 1], a);
    }
    proj_set_zero(c);
    bool found = false;
    for (int nib = kLimbs * 8 - 1; nib >= 0; nib--) {
      const uint32_t d = (k.l[nib / 8] >> ((nib % 8) * 4)) & 15u;
      if (found)
        for (int t = 0; t < 4; t++) proj_dbl<G>(c, c);
      if (d) {
        if (found) proj_add<G>(c, c, tab[d]);
        else c = tab[d];
        found = true;
      }
    }
    memcpy(out, &c, sizeof(c));
  }
  static void zero(void *out) {  // O = (0 : 1 : 0)
    Proj<F> z;
    proj_set_zero(z);
    memcpy(out, &z, sizeof(z));
  }
  static void to_affine(const void *p, void *out) {
    Proj<F> a;
    Affine<F> o;
    memcpy(&a, p, sizeof(a));
    proj_to_affine<G>(o, a);
    memcpy(out, &o, sizeof(o));
  }
  // 2^k * P through the Jacobian doubling of the base-table builder, affine wire format (test hook: the same template
  // runs on the device)
  static void jacobian_doublings(const void *xy, int k, void *out) {
    Affine<F> a, o;
    memcpy(&a, xy, sizeof(a));
    Proj<F> cur;
    cur.X = a.x;
    cur.Y = a.y;
    F::set_one(cur.Z);
    for (int i = 0; i < k; i++) jac_dbl<G>(cur, cur);
    F zi, zi2;
    F::inv(zi, cur.Z);
    F::sqr(zi2, zi);
    F::mul(o.x, cur.X, zi2);
    F::mul(zi2, zi2, zi);
    F::mul(o.y, cur.Y, zi2);
    memcpy(out, &o, sizeof(o));
  }
  static void from_affine(const void *xy, void *out) {
    Affine<F> a;
    Proj<F> p;
    memcpy(&a, xy, sizeof(a));
    proj_from_affine(p, a);
    memcpy(out, &p, sizeof(p));
  }
};

#define DISPATCH_GROUP(curve, group, CALL)                                  \
  do {                                                                      \
    if ((curve) == 0 && (group) == 1) { HostOps<Mnt4G1>::CALL; return 0; }  \
    if ((curve) == 0 && (group) == 2) { HostOps<Mnt4G2>::CALL; return 0; }  \
    if ((curve) == 1 && (group) == 1) { HostOps<Mnt6G1>::CALL; return 0; }  \
    if ((curve) == 1 && (group) == 2) { HostOps<Mnt6G2>::CALL; return 0; }  \
    return set_error(-1, "bad curve %d", (int)(curve));                     \
  } while (0)

static int group_zero(int curve, int group, void *out) { DISPATCH_GROUP(curve, group, zero(out)); }
static size_t g2_degree(int curve) { return curve == 0 ? 2 : 3; }
static size_t affine_bytes(int curve, int group) { return 2 * 96 * (group == 1 ? 1 : g2_degree(curve)); }
static size_t proj_bytes(int curve, int group) { return 3 * 96 * (group == 1 ? 1 : g2_degree(curve)); }

static int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return set_error(-10, "no usable CUDA device (%s); this library has no CPU fallback",
                     e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  return 0;
}
}  // namespace b200

using namespace b200;

template <class P>
static void host_fp_op_t(int op, const void *a, const void *b, void *r) {
  Fp<P> x, y, z;
  memcpy(&x, a, 96);
  if (b) memcpy(&y, b, 96);
  else Fp<P>::set_zero(y);
  switch (op) {
    case 0: Fp<P>::add(z, x, y); break;
    case 1: Fp<P>::sub(z, x, y); break;
    case 2: Fp<P>::mul(z, x, y); break;
    case 3: Fp<P>::inv(z, x); break;
    case 4: Fp<P>::from_mont(z, x); break;
    case 6: Fp<P>::inv_binary(z, x); break;
    case 7: Fp<P>::inv_bingcd(z, x); break;
    case 8: Fp<P>::inv_fermat(z, x); break;
    default: Fp<P>::to_mont(z, x); break;
  }
  memcpy(r, &z, 96);
}

struct b200_domain {
  Domain *impl;
};

// Pre-shifted base tables for one slicing of the five queries (see msm_precompute_kernel).
struct Precomputed {
  // the slicing the tables were built for: query qi (0 A, 1 B1, 2 B2, 3 L, 4 H) covers slices [spans[qi][0], spans[qi][1])
  // of `world` (an empty span: no table)
  int spans[5][2] = {{-1, -1}, {-1, -1}, {-1, -1}, {-1, -1}, {-1, -1}}, world = -1;
  DevBuf table[5];
  MsmPlan plan[5];
  MsmDedup dedup[5];  // equal bases inside this rank's slice of each G1 query (job order A, B1, B2, H, L)
  double build_ms = 0;
};

struct b200_params {
  int curve;
  size_t d, m;
  double load_ms[3] = {0, 0, 0};  // from_file: read, copy wait, total
  const void *q[5];  // A, B1, B2, L, H (device)
  DevBuf owned;      // backing store when loaded from a host image
  b200_domain *dom;
  DevBuf w, ca, cb, cc, h;  // per-proof device buffers
  Precomputed pre;
};

// B200_SHARE_PREP=0: every MSM of a proof extracts and sorts its own digits (A/B timing and tests of the sharing)
static bool share_prep_enabled() {
  static const bool on = !(getenv("B200_SHARE_PREP") && getenv("B200_SHARE_PREP")[0] == '0');
  return on;
}
static std::atomic<int> g_use_precompute{-1};  // -1: read B200_PRECOMPUTE from the environment (default on)
static bool use_precompute() {
  if (g_use_precompute < 0) {
    const char *e = getenv("B200_PRECOMPUTE");
    g_use_precompute = (e && e[0] == '0') ? 0 : 1;
  }
  return g_use_precompute != 0;
}

// Equal bases of one G1 point range (key-load time, host): 64-bit hash of the 192 wire bytes, sort, confirm with memcmp.
// Only worth a separate scalar pass when a sizeable part of the range folds away (the A query: m/2 of m+1 bases);
// a stray duplicate pair (B1, B2) is left to the P+P branch of the bucket accumulation.
// host part: groups of equal points in `pts` (n wire-format points). members: indices grouped, representative first;
// segments: (first position in members, length, group) triples; groups: (representative, first segment, #segments)
// triples. Returns the number of bases folded away. Points at infinity (y == 0) are never grouped: the kernels skip them.
static size_t group_equal_bases(const unsigned char *pts, size_t n, size_t point_bytes, std::vector<uint32_t> &members,
                                std::vector<uint32_t> &segments, std::vector<uint32_t> &groups) {
  members.clear();
  segments.clear();
  groups.clear();
  std::vector<std::pair<uint64_t, uint32_t>> keyed;
  keyed.reserve(n);
  for (size_t i = 0; i < n; i++) {
    const unsigned char *q = pts + i * point_bytes;
    bool inf = true;
    for (size_t k = point_bytes / 2; k < point_bytes && inf; k++) inf = q[k] == 0;
    if (inf) continue;
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t k = 0; k < point_bytes; k += 8) {
      uint64_t w;
      memcpy(&w, q + k, 8);
      h = (h ^ w) * 0x100000001b3ull;
      h ^= h >> 29;
    }
    keyed.emplace_back(h, (uint32_t)i);
  }
  std::sort(keyed.begin(), keyed.end());
  size_t merged = 0;
  for (size_t a = 0; a < keyed.size();) {
    size_t b = a + 1;
    while (b < keyed.size() && keyed[b].first == keyed[a].first) b++;
    if (b - a >= 2) {
      // members of the run that really equal its first element (a hash collision just stays unmerged)
      const unsigned char *rep = pts + (size_t)keyed[a].second * point_bytes;
      const size_t first = members.size();
      members.push_back(keyed[a].second);
      for (size_t k = a + 1; k < b; k++)
        if (memcmp(rep, pts + (size_t)keyed[k].second * point_bytes, point_bytes) == 0)
          members.push_back(keyed[k].second);
      const size_t len = members.size() - first;
      if (len < 2) {
        members.resize(first);
      } else {
        const uint32_t g = (uint32_t)(groups.size() / 3), seg0 = (uint32_t)(segments.size() / 3);
        for (size_t o = 0; o < len; o += kDedupSegment) {
          segments.push_back((uint32_t)(first + o));
          segments.push_back((uint32_t)std::min<size_t>(kDedupSegment, len - o));
          segments.push_back(g);
        }
        groups.push_back(keyed[a].second);
        groups.push_back(seg0);
        groups.push_back((uint32_t)(segments.size() / 3) - seg0);
        merged += len - 1;
      }
    }
    a = b;
  }
  return merged;
}

// (copies go through `st`, a non-blocking stream: the caller runs this on a host thread of its own while the table
// kernels occupy the default stream)
static int find_equal_bases(const void *d_points, size_t n, size_t point_bytes, MsmDedup &dd, cudaStream_t st) {
  dd.reset();
  if (n < 8) return 0;
  std::vector<unsigned char> pts(n * point_bytes);
  B200_CUDA_CHECK(cudaMemcpyAsync(pts.data(), d_points, pts.size(), cudaMemcpyDeviceToHost, st));
  B200_CUDA_CHECK(cudaStreamSynchronize(st));
  std::vector<uint32_t> members, segments, groups;
  const size_t merged = group_equal_bases(pts.data(), n, point_bytes, members, segments, groups);
  if (merged < std::max<size_t>(4, n / 32)) return 0;
  dd.nsegments = (uint32_t)(segments.size() / 3);
  dd.ngroups = (uint32_t)(groups.size() / 3);
  B200_CHECK(dd.members.reserve(members.size() * 4));
  B200_CHECK(dd.segments.reserve(segments.size() * 4));
  B200_CHECK(dd.groups.reserve(groups.size() * 4));
  B200_CHECK(dd.segment_sums.reserve((size_t)dd.nsegments * 96));
  B200_CUDA_CHECK(cudaMemcpyAsync(dd.members.p, members.data(), members.size() * 4, cudaMemcpyHostToDevice, st));
  B200_CUDA_CHECK(cudaMemcpyAsync(dd.segments.p, segments.data(), segments.size() * 4, cudaMemcpyHostToDevice, st));
  B200_CUDA_CHECK(cudaMemcpyAsync(dd.groups.p, groups.data(), groups.size() * 4, cudaMemcpyHostToDevice, st));
  B200_CUDA_CHECK(cudaStreamSynchronize(st));
  dd.merged = merged;
  return 0;
}

extern "C" {

const char *b200_version(void) { return "b200-groth16-mnt753 0.1 (sm_100a)"; }
const char *b200_last_error(void) { return last_error().c_str(); }
int b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}
int b200_set_device(int ordinal) {
  B200_CHECK(require_device());
  B200_CUDA_CHECK(cudaSetDevice(ordinal));
  return 0;
}
int b200_sync(void) {
  B200_CUDA_CHECK(cudaDeviceSynchronize());
  return 0;
}
int b200_malloc(void **d_ptr, size_t bytes) {
  B200_CHECK(require_device());
  B200_CUDA_CHECK(cudaMalloc(d_ptr, bytes ? bytes : 16));
  return 0;
}
int b200_free(void *d_ptr) {
  B200_CUDA_CHECK(cudaFree(d_ptr));
  return 0;
}
int b200_host_alloc(void **h_ptr, size_t bytes) {
  B200_CHECK(require_device());
  B200_CUDA_CHECK(cudaMallocHost(h_ptr, bytes ? bytes : 16));
  return 0;
}
int b200_host_free(void *h_ptr) {
  B200_CUDA_CHECK(cudaFreeHost(h_ptr));
  return 0;
}
int b200_memcpy_h2d(void *d, const void *h, size_t bytes) {
  B200_CUDA_CHECK(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice));
  return 0;
}
int b200_memcpy_d2h(void *h, const void *d, size_t bytes) {
  B200_CUDA_CHECK(cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost));
  return 0;
}
int b200_memcpy_d2d(void *d, const void *s, size_t bytes) {
  B200_CUDA_CHECK(cudaMemcpy(d, s, bytes, cudaMemcpyDeviceToDevice));
  return 0;
}
int b200_memset_zero(void *d, size_t bytes) {
  B200_CUDA_CHECK(cudaMemset(d, 0, bytes));
  return 0;
}

// ---- Fr vectors
int b200_fr_muleq(int curve, void *d_a, const void *d_b, size_t n) {
  B200_CHECK(require_device());
  return fr_muleq(curve, d_a, d_b, n);
}
int b200_fr_subeq(int curve, void *d_a, const void *d_b, size_t n) {
  B200_CHECK(require_device());
  return fr_subeq(curve, d_a, d_b, n);
}

// ---- domain
int b200_domain_create(int curve, size_t m, b200_domain **out) {
  B200_CHECK(require_device());
  if (curve != 0 && curve != 1) return set_error(-1, "bad curve %d", curve);
  Domain *impl = nullptr;
  B200_CHECK(domain_create(curve, m, &impl));
  *out = new b200_domain{impl};
  return 0;
}
int b200_domain_destroy(b200_domain *dom) {
  if (dom) {
    domain_destroy(dom->impl);
    delete dom;
  }
  return 0;
}
size_t b200_domain_size(const b200_domain *dom) { return domain_size(dom->impl); }
int b200_domain_table(const b200_domain *dom, int which, void *h_out, size_t count) {
  return domain_table(dom->impl, which, h_out, count);
}
int b200_domain_fft(b200_domain *dom, void *d_a) { return domain_transform(dom->impl, d_a, 0); }
int b200_domain_ifft(b200_domain *dom, void *d_a) { return domain_transform(dom->impl, d_a, 1); }
int b200_domain_coset_fft(b200_domain *dom, void *d_a) { return domain_transform(dom->impl, d_a, 2); }
int b200_domain_icoset_fft(b200_domain *dom, void *d_a) { return domain_transform(dom->impl, d_a, 3); }
int b200_domain_divide_by_z_on_coset(b200_domain *dom, void *d_a) { return domain_divide_by_z(dom->impl, d_a); }
int b200_compute_h(b200_domain *dom, void *d_ca, void *d_cb, void *d_cc, void *d_out) {
  return compute_h(dom->impl, d_ca, d_cb, d_cc, d_out);
}

// ---- MSM
int b200_msm_g1(int curve, const void *d_scalars, const void *d_points, size_t n, void *h_out) {
  B200_CHECK(require_device());
  return msm_dispatch(curve, 1, d_scalars, d_points, n, h_out);
}
int b200_msm_g2(int curve, const void *d_scalars, const void *d_points, size_t n, void *h_out) {
  B200_CHECK(require_device());
  return msm_dispatch(curve, 2, d_scalars, d_points, n, h_out);
}
struct b200_msm_ctx {
  int curve, group;
  size_t n;
  DevBuf table;
  MsmPlan plan;
};
int b200_msm_ctx_create(int curve, int group, const void *d_points, size_t n, b200_msm_ctx **out) {
  B200_CHECK(require_device());
  if ((curve != 0 && curve != 1) || (group != 1 && group != 2)) return set_error(-1, "bad curve/group %d/%d", curve, group);
  if (n == 0 || !d_points) return set_error(-1, "msm_ctx_create: empty point set");
  std::unique_ptr<b200_msm_ctx> c(new b200_msm_ctx());
  c->curve = curve;
  c->group = group;
  c->n = n;
  B200_CHECK(msm_precompute_dispatch(curve, group, d_points, n, c->plan, c->table));
  *out = c.release();
  return 0;
}
int b200_msm_ctx_run(b200_msm_ctx *ctx, const void *d_scalars, void *h_out) {
  B200_CHECK(require_device());
  MsmTail tail;
  msm_select_slot(0);
  B200_CHECK(msm_table_dispatch_deferred(ctx->curve, ctx->group, d_scalars, ctx->table.p, ctx->n, ctx->plan, h_out, tail,
                                         MsmShare(), nullptr));
  std::string err;
  const int rc = tail(err);
  if (rc) return set_error(rc, "%s", err.c_str());
  return 0;
}
int b200_msm_ctx_destroy(b200_msm_ctx *ctx) {
  delete ctx;
  return 0;
}
int b200_msm_set_batch_affine(int on) {
  msm_set_batch_affine(on);
  return 0;
}
int b200_msm_get_batch_affine(void) { return msm_accum_mode(); }
int b200_msm_set_window(int c) {
  msm_set_window(c);
  return 0;
}
unsigned long long b200_launch_count(void) { return launch_counter(); }
int b200_msm_phase_totals(double *out5, int reset) {
  msm_phase_totals(out5, reset);
  return 0;
}
int b200_msm_last_plan(int *out3) {
  msm_last_plan(out3);
  return 0;
}
int b200_host_equal_bases(const void *h_points, size_t n, size_t point_bytes, uint32_t *members, size_t *n_members,
                          uint32_t *groups, size_t *n_groups, size_t *merged) {
  if (!h_points || point_bytes == 0 || point_bytes % 16) return set_error(-1, "equal_bases: bad arguments");
  std::vector<uint32_t> m, s, g;
  const size_t folded = group_equal_bases((const unsigned char *)h_points, n, point_bytes, m, s, g);
  if (members && n_members && *n_members >= m.size()) memcpy(members, m.data(), m.size() * 4);
  if (groups && n_groups && *n_groups >= g.size()) memcpy(groups, g.data(), g.size() * 4);
  if (n_members) *n_members = m.size();
  if (n_groups) *n_groups = g.size();
  if (merged) *merged = folded;
  return 0;
}
int b200_prove_timeline(int begin, double *out15) {
  if (begin) msm_timeline_begin();
  else if (out15) msm_timeline_get(out15);
  return 0;
}
int b200_msm_last_phase_ms(double *out5) {
  msm_last_phase_ms(out5);
  return 0;
}

// ---- host helpers
int b200_g1_add(int curve, const void *p, const void *q, void *out) { DISPATCH_GROUP(curve, 1, add(p, q, out)); }
int b200_g2_add(int curve, const void *p, const void *q, void *out) { DISPATCH_GROUP(curve, 2, add(p, q, out)); }
int b200_g1_scale(int curve, const void *fr, const void *p, void *out) { DISPATCH_GROUP(curve, 1, scale(fr, p, out)); }
int b200_g2_scale(int curve, const void *fr, const void *p, void *out) { DISPATCH_GROUP(curve, 2, scale(fr, p, out)); }
int b200_g1_to_affine(int curve, const void *p, void *out) { DISPATCH_GROUP(curve, 1, to_affine(p, out)); }
int b200_g2_to_affine(int curve, const void *p, void *out) { DISPATCH_GROUP(curve, 2, to_affine(p, out)); }
int b200_g1_from_affine(int curve, const void *xy, void *out) { DISPATCH_GROUP(curve, 1, from_affine(xy, out)); }
int b200_g2_from_affine(int curve, const void *xy, void *out) { DISPATCH_GROUP(curve, 2, from_affine(xy, out)); }

int b200_host_jacobian_doublings(int curve, int group, const void *h_xy, int k, void *h_out_xy) {
  if (k < 0 || k > 4096) return set_error(-1, "bad doubling count");
  DISPATCH_GROUP(curve, group, jacobian_doublings(h_xy, k, h_out_xy));
}
int b200_host_fp_op(int tag, int op, const void *a, const void *b, void *r) {
  if (op < 0 || op > 8) return set_error(-1, "bad op");
  if (tag == 0) host_fp_op_t<PrimeA>(op, a, b, r);
  else host_fp_op_t<PrimeB>(op, a, b, r);
  return 0;
}

// ---- test hooks
int b200_dev_fp_op(int tag, int op, const void *a, const void *b, void *r, size_t n) {
  B200_CHECK(require_device());
  return dev_fp_op(tag, op, a, b, r, n);
}
int b200_dev_fqe_op(int curve, int op, const void *a, const void *b, void *r, size_t n) {
  B200_CHECK(require_device());
  return dev_fqe_op(curve, op, a, b, r, n);
}
int b200_dev_group_op(int curve, int group, int op, const void *p, const void *q, void *r, size_t n) {
  B200_CHECK(require_device());
  return dev_group_op(curve, group, op, p, q, r, n);
}
int b200_gen_points(int curve, int group, void *d_out, size_t n, uint64_t first) {
  B200_CHECK(require_device());
  return gen_points(curve, group, d_out, n, first);
}
int b200_batch_exp(int curve, int group, const void *h_base_affine, const void *d_scalars, size_t n, void *d_out_affine,
                   int window, double *ms3) {
  B200_CHECK(require_device());
  if (curve != 0 && curve != 1) return set_error(-1, "bad curve %d", curve);
  if (group != 1 && group != 2) return set_error(-1, "bad group %d", group);
  if (!h_base_affine || (n && (!d_scalars || !d_out_affine))) return set_error(-1, "batch_exp: null argument");
  return batch_exp(curve, group, h_base_affine, d_scalars, n, d_out_affine, window, ms3);
}
int b200_imad_peak(double *mac32_per_s, double *ms) {
  B200_CHECK(require_device());
  return imad_peak(mac32_per_s, ms);
}

// ---- proving key
// Dimensions come from a file header or from the caller: validate them before any size arithmetic (m - 1 and the byte
// counts below would wrap). The query lengths are m+1, m+1, m+1, m-1, d; the domain has d+1 elements.
static int check_key_dims(size_t d, size_t m) {
  const size_t kMax = (size_t)1 << 30;
  if (m < 2 || m >= kMax || d < 1 || d >= kMax)
    return set_error(-4, "key dimensions d=%zu m=%zu out of range (need 1 <= d < 2^30, 2 <= m < 2^30)", d, m);
  return 0;
}
static void params_set_queries(b200_params *p) {
  const size_t g1 = affine_bytes(p->curve, 1), g2 = affine_bytes(p->curve, 2), m = p->m;
  char *base = (char *)p->owned.p;
  p->q[0] = base;
  p->q[1] = base + g1 * (m + 1);
  p->q[2] = base + 2 * g1 * (m + 1);
  p->q[3] = base + 2 * g1 * (m + 1) + g2 * (m + 1);
  p->q[4] = base + 2 * g1 * (m + 1) + g2 * (m + 1) + g1 * (m - 1);
}

static int params_finish(b200_params *p) {
  b200_domain *dom = nullptr;
  B200_CHECK(b200_domain_create(p->curve, p->d + 1, &dom));
  p->dom = dom;
  B200_CHECK(p->w.alloc((p->m + 1) * 96));
  B200_CHECK(p->ca.alloc((p->d + 1) * 96));
  B200_CHECK(p->cb.alloc((p->d + 1) * 96));
  B200_CHECK(p->cc.alloc((p->d + 1) * 96));
  B200_CHECK(p->h.alloc((p->d + 2) * 96));
  return 0;
}
int b200_params_from_host(int curve, const void *h_image, size_t bytes, b200_params **out) {
  B200_CHECK(require_device());
  if (curve != 0 && curve != 1) return set_error(-1, "bad curve %d", curve);
  if (bytes < 16) return set_error(-4, "parameter image too short");
  size_t d, m;
  memcpy(&d, h_image, 8);
  memcpy(&m, (const char *)h_image + 8, 8);
  B200_CHECK(check_key_dims(d, m));
  size_t g1 = affine_bytes(curve, 1), g2 = affine_bytes(curve, 2);
  size_t need = 16 + g1 * (2 * (m + 1) + (m - 1) + d) + g2 * (m + 1);
  if (bytes != need)
    return set_error(-4, "parameter image has %zu bytes, expected %zu for d=%zu m=%zu", bytes, need, d, m);
  b200_params *p = new b200_params();
  p->curve = curve;
  p->d = d;
  p->m = m;
  p->dom = nullptr;
  int rc = p->owned.alloc(bytes - 16);
  if (rc) {
    delete p;
    return rc;
  }
  cudaError_t e = cudaMemcpy(p->owned.p, (const char *)h_image + 16, bytes - 16, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    delete p;
    return set_error(-100 - (int)e, "H2D of parameters failed: %s", cudaGetErrorString(e));
  }
  params_set_queries(p);
  rc = params_finish(p);
  if (rc) {
    b200_params_destroy(p);
    return rc;
  }
  *out = p;
  return 0;
}
// file (positioned at the first byte to copy) -> device, `bytes` bytes: two pinned staging buffers, chunk k's
// host->device copy runs while chunk k+1 is being read. The buffers, their events and the copy stream are made once per
// device and kept (pinning 128 MB costs ~100 ms - more than copying a 100 MB vector).
namespace {
struct StagePool {
  static constexpr size_t kChunk = (size_t)64 << 20;
  void *h[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  cudaStream_t stream = nullptr;
  int device = -1;
  std::mutex mu;
  int ensure() {
    int dev = 0;
    B200_CUDA_CHECK(cudaGetDevice(&dev));
    if (device == dev) return 0;
    release();
    B200_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      B200_CUDA_CHECK(cudaMallocHost(&h[i], kChunk));
      B200_CUDA_CHECK(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    }
    device = dev;
    return 0;
  }
  void release() {
    for (int i = 0; i < 2; i++) {
      if (h[i]) cudaFreeHost(h[i]);
      if (done[i]) cudaEventDestroy(done[i]);
      h[i] = nullptr;
      done[i] = nullptr;
    }
    if (stream) cudaStreamDestroy(stream);
    stream = nullptr;
    device = -1;
  }
};
StagePool &stage_pool() {
  static StagePool *pool = new StagePool();  // leaked on purpose: outlives static destruction of the CUDA runtime
  return *pool;
}
}  // namespace
static int stream_file_to_device(FILE *f, const char *path, void *d_dst, size_t bytes, double &read_ms, double &wait_ms) {
  StagePool &sp = stage_pool();
  std::lock_guard<std::mutex> g(sp.mu);
  B200_CHECK(sp.ensure());
  const size_t chunk = StagePool::kChunk;
  int k = 0;
  for (size_t off = 0; off < bytes; off += chunk, k ^= 1) {
    const size_t n = bytes - off < chunk ? bytes - off : chunk;
    double a = now_ms();
    B200_CUDA_CHECK(cudaEventSynchronize(sp.done[k]));  // the copy that last used this buffer (no-op the first time)
    double b = now_ms();
    if (fread(sp.h[k], 1, n, f) != n) return set_error(-4, "short read on %s", path);
    double c = now_ms();
    wait_ms += b - a;
    read_ms += c - b;
    B200_CUDA_CHECK(cudaMemcpyAsync((char *)d_dst + off, sp.h[k], n, cudaMemcpyHostToDevice, sp.stream));
    B200_CUDA_CHECK(cudaEventRecord(sp.done[k], sp.stream));
  }
  double a = now_ms();
  B200_CUDA_CHECK(cudaStreamSynchronize(sp.stream));
  wait_ms += now_ms() - a;
  return 0;
}

int b200_file_to_device(const char *path, size_t file_offset, void *d_dst, size_t bytes) {
  B200_CHECK(require_device());
  FILE *f = fopen(path, "rb");
  if (!f) return set_error(-4, "cannot open %s", path);
  struct Closer {
    FILE *f;
    ~Closer() { fclose(f); }
  } closer{f};
  if (fseek(f, (long)file_offset, SEEK_SET) != 0) return set_error(-4, "cannot seek in %s", path);
  double r = 0, w = 0;
  return stream_file_to_device(f, path, d_dst, bytes, r, w);
}

b200_domain *b200_params_domain(const b200_params *p) { return p ? p->dom : nullptr; }

int b200_params_from_file(int curve, const char *path, b200_params **out) {
  B200_CHECK(require_device());
  if (curve != 0 && curve != 1) return set_error(-1, "bad curve %d", curve);
  const double t0 = now_ms();
  FILE *f = fopen(path, "rb");
  if (!f) return set_error(-4, "cannot open parameter file %s", path);
  struct Closer {
    FILE *f;
    ~Closer() { fclose(f); }
  } closer{f};
  uint64_t hdr[2];
  if (fread(hdr, 1, 16, f) != 16) return set_error(-4, "parameter file %s is shorter than its header", path);
  const size_t d = (size_t)hdr[0], m = (size_t)hdr[1];
  B200_CHECK(check_key_dims(d, m));
  const size_t g1 = affine_bytes(curve, 1), g2 = affine_bytes(curve, 2);
  const size_t body = g1 * (2 * (m + 1) + (m - 1) + d) + g2 * (m + 1);
  if (fseek(f, 0, SEEK_END) != 0) return set_error(-4, "cannot seek in %s", path);
  const long fsize = ftell(f);
  if (fsize < 0 || (size_t)fsize != body + 16)
    return set_error(-4, "parameter file %s has %ld bytes, expected %zu for d=%zu m=%zu", path, fsize, body + 16, d, m);
  fseek(f, 16, SEEK_SET);
  std::unique_ptr<b200_params> p(new b200_params());
  p->curve = curve;
  p->d = d;
  p->m = m;
  p->dom = nullptr;
  B200_CHECK(p->owned.alloc(body));
  double read_ms = 0, wait_ms = 0;
  B200_CHECK(stream_file_to_device(f, path, p->owned.p, body, read_ms, wait_ms));
  params_set_queries(p.get());
  int rc = params_finish(p.get());
  if (rc) {
    b200_domain_destroy(p->dom);
    return rc;
  }
  p->load_ms[0] = read_ms;
  p->load_ms[1] = wait_ms;
  p->load_ms[2] = now_ms() - t0;
  *out = p.release();
  return 0;
}
int b200_params_load_ms(const b200_params *p, double *out3) {
  for (int i = 0; i < 3; i++) out3[i] = p->load_ms[i];
  return 0;
}

int b200_params_from_device(int curve, size_t d, size_t m, const void *A, const void *B1, const void *B2, const void *L,
                            const void *H, b200_params **out) {
  B200_CHECK(require_device());
  if (curve != 0 && curve != 1) return set_error(-1, "bad curve %d", curve);
  B200_CHECK(check_key_dims(d, m));
  if (!A || !B1 || !B2 || !L || !H) return set_error(-1, "params_from_device: null query pointer");
  b200_params *p = new b200_params();
  p->curve = curve;
  p->d = d;
  p->m = m;
  p->dom = nullptr;
  p->q[0] = A; p->q[1] = B1; p->q[2] = B2; p->q[3] = L; p->q[4] = H;
  int rc = params_finish(p);
  if (rc) {
    b200_params_destroy(p);
    return rc;
  }
  *out = p;
  return 0;
}
int b200_params_destroy(b200_params *p) {
  if (p) {
    b200_domain_destroy(p->dom);
    delete p;
  }
  return 0;
}
size_t b200_params_d(const b200_params *p) { return p->d; }
size_t b200_params_m(const b200_params *p) { return p->m; }
const void *b200_params_query(const b200_params *p, int which) { return (which >= 0 && which < 5) ? p->q[which] : nullptr; }

// Slice [lo, hi) of query qi (0 A, 1 B1, 2 B2, 3 L, 4 H) that rank `rank` of `world` sums: contiguous ranges like
// multi_exp's chunks (multiexp.tcc:417-431; the last rank takes the remainder). The four w-driven queries are cut so
// that a rank's scalars are ONE range of w - A / B1 / B2 take points [lo1, hi1) of m+1, and L, whose point i belongs to
// w[i + 2] (main.cpp:247-250), takes points [lo1 - 2, hi1 - 2) clipped to [0, m-1) - so the four MSMs of a rank share
// one digit extraction and one counting sort at every world size.
// A caller may own a RUN of consecutive slices, [rank, rank_end) of `world` (uneven sharding: bench.py gives the GPU
// that also proves the small curve a smaller share of the large one); its range is their union.
static void query_slice(size_t d, size_t m, int qi, int rank, int rank_end, int world, size_t &lo, size_t &hi) {
  auto cut = [&](size_t n, int r) -> size_t { return r >= world ? n : (size_t)r * (n / (size_t)world); };
  if (qi == 4) {
    lo = cut(d, rank);
    hi = cut(d, rank_end);
    return;
  }
  const size_t n1 = m + 1;
  const size_t lo1 = cut(n1, rank), hi1 = cut(n1, rank_end);
  if (qi != 3) {
    lo = lo1;
    hi = hi1;
    return;
  }
  lo = lo1 >= 2 ? lo1 - 2 : 0;
  hi = hi1 >= 2 ? hi1 - 2 : 0;
  if (hi > m - 1) hi = m - 1;
  if (lo > hi) lo = hi;
}

// Build (once per key and slicing) the tables of pre-shifted bases for this rank's slice of the five queries.
int b200_params_precompute(b200_params *p, int rank, int world) { return b200_params_precompute_span(p, rank, rank + 1, world); }
int b200_params_precompute_span(b200_params *p, int rank, int rank_end, int world) {
  const int spans[10] = {rank, rank_end, rank, rank_end, rank, rank_end, rank, rank_end, rank, rank_end};
  return b200_params_precompute_queries(p, spans, world);
}
static int check_spans(const int *spans, int world) {
  if (!spans || world < 1) return set_error(-1, "bad slicing (world %d)", world);
  for (int q = 0; q < 5; q++)
    if (spans[2 * q] < 0 || spans[2 * q + 1] < spans[2 * q] || spans[2 * q + 1] > world)
      return set_error(-1, "bad slice run [%d, %d) of %d for query %d", spans[2 * q], spans[2 * q + 1], world, q);
  return 0;
}
int b200_params_precompute_queries(b200_params *p, const int *spans, int world) {
  B200_CHECK(require_device());
  B200_CHECK(check_spans(spans, world));
  if (p->pre.world == world && memcmp(p->pre.spans, spans, sizeof(p->pre.spans)) == 0) return 0;
  double t0 = now_ms();
  const size_t d = p->d, m = p->m;
  const int jobq[5] = {0, 1, 2, 4, 3};                     // job order A, B1, B2, H, L -> query index
  p->pre.world = -1;
  // The grouping of equal bases (host: hash + sort of the G1 queries' wire bytes) runs on a thread of its own, under
  // the table kernels.
  int dev = 0;
  B200_CUDA_CHECK(cudaGetDevice(&dev));
  struct Slice { const char *pts; size_t n; int group; };
  Slice slices[5];
  for (int j = 0; j < 5; j++) {
    const int qi = jobq[j];
    size_t lo, hi;
    query_slice(d, m, qi, spans[2 * qi], spans[2 * qi + 1], world, lo, hi);
    slices[j].group = qi == 2 ? 2 : 1;
    slices[j].pts = (const char *)p->q[qi] + lo * affine_bytes(p->curve, slices[j].group);
    slices[j].n = hi - lo;
    p->pre.dedup[j].reset();
  }
  int dedup_rc = 0;
  std::string dedup_err;
  std::thread dedup_thread([&] {
    cudaStream_t st = nullptr;
    if (cudaSetDevice(dev) != cudaSuccess || cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
      dedup_rc = set_error(-2, "equal-base grouping: cannot set up device %d", dev);
      dedup_err = last_error();
      return;
    }
    for (int j = 0; j < 5 && dedup_rc == 0; j++)
      if (slices[j].n > 0 && slices[j].group == 1) {
        dedup_rc = find_equal_bases(slices[j].pts, slices[j].n, affine_bytes(p->curve, 1), p->pre.dedup[j], st);
        if (dedup_rc) dedup_err = last_error();
      }
    cudaStreamDestroy(st);
  });
  int rc = 0;
  const bool verbose = getenv("B200_VERBOSE") != nullptr;
  for (int j = 0; j < 5 && rc == 0; j++) {
    p->pre.plan[j] = MsmPlan();
    const double a = now_ms();
    if (slices[j].n > 0)
      rc = msm_precompute_dispatch(p->curve, slices[j].group, slices[j].pts, slices[j].n, p->pre.plan[j], p->pre.table[j]);
    if (verbose) fprintf(stderr, "[b200] base table %d (G%d, %zu points, %d windows): %.0f ms\n", j, slices[j].group, slices[j].n, p->pre.plan[j].W, now_ms() - a);
  }
  const double tj = now_ms();
  dedup_thread.join();
  if (verbose) fprintf(stderr, "[b200] waited %.0f ms for the equal-base grouping\n", now_ms() - tj);
  if (rc) return rc;
  if (dedup_rc) return set_error(dedup_rc, "%s", dedup_err.c_str());
  memcpy(p->pre.spans, spans, sizeof(p->pre.spans));
  p->pre.world = world;
  p->pre.build_ms = now_ms() - t0;
  return 0;
}
double b200_params_precompute_ms(const b200_params *p) { return p->pre.build_ms; }
int b200_set_precompute(int on) {
  g_use_precompute = on ? 1 : 0;
  return 0;
}

// MSM of one whole query of a resident key: uses the pre-shifted base table when one exists for (rank 0, world 1)
// and n is the query's length, else the table-free path. This is what B::multiexp_G1/G2 call, so the reference's own
// driver gets the table speed-up too.
int b200_params_msm(b200_params *p, int which, const void *d_scalars, size_t n, void *h_out) {
  B200_CHECK(require_device());
  if (which < 0 || which > 4) return set_error(-1, "bad query index %d", which);
  const size_t ns[5] = {p->m + 1, p->m + 1, p->m + 1, p->m - 1, p->d};  // A, B1, B2, L, H
  const int job_of_query[5] = {0, 1, 2, 4, 3};
  const int group = which == 2 ? 2 : 1;
  if (use_precompute() && n == ns[which]) {
    B200_CHECK(b200_params_precompute(p, 0, 1));
    const int j = job_of_query[which];
    MsmTail tail;
    msm_select_slot(0);
    B200_CHECK(msm_table_dispatch_deferred(p->curve, group, d_scalars, p->pre.table[j].p, n, p->pre.plan[j], h_out, tail,
                                           MsmShare(), &p->pre.dedup[j]));
    std::string err;
    const int rc = tail(err);
    if (rc) return set_error(rc, "%s", err.c_str());
    return 0;
  }
  return msm_dispatch(p->curve, group, d_scalars, p->q[which], n, h_out);
}

struct b200_msm_pending {
  MsmTail tail;
};
int b200_params_msm_async(b200_params *p, int which, const void *d_scalars, size_t n, void *h_out, b200_msm_pending **out) {
  B200_CHECK(require_device());
  if (which < 0 || which > 4) return set_error(-1, "bad query index %d", which);
  const size_t ns[5] = {p->m + 1, p->m + 1, p->m + 1, p->m - 1, p->d};  // A, B1, B2, L, H
  const int job_of_query[5] = {0, 1, 2, 4, 3};
  const int group = which == 2 ? 2 : 1;
  static thread_local int next_slot = 0;  // round robin over the calling thread's five workspaces
  msm_select_slot(next_slot);
  next_slot = (next_slot + 1) % kMsmSlots;
  std::unique_ptr<b200_msm_pending> h(new b200_msm_pending());
  int rc;
  if (use_precompute() && n == ns[which]) {
    B200_CHECK(b200_params_precompute(p, 0, 1));
    const int j = job_of_query[which];
    rc = msm_table_dispatch_deferred(p->curve, group, d_scalars, p->pre.table[j].p, n, p->pre.plan[j], h_out, h->tail, MsmShare(),
                                     &p->pre.dedup[j]);
  } else {
    rc = msm_dispatch_deferred(p->curve, group, d_scalars, p->q[which], n, h_out, h->tail);
  }
  msm_select_slot(0);
  if (rc) return rc;
  *out = h.release();
  return 0;
}
int b200_msm_wait(b200_msm_pending *pending) {
  if (!pending) return 0;
  std::unique_ptr<b200_msm_pending> h(pending);
  std::string err;
  const int rc = h->tail(err);
  if (rc) return set_error(rc, "%s", err.c_str());
  return 0;
}

// Called once by prove_partials on the issuing thread when all of a proof's MSMs have been issued (b200_prove_batch can
// start the next proof of a batch at that moment instead of at time zero: B200_BATCH_STAGGER).
static thread_local std::function<void()> *tl_all_issued_hook = nullptr;

// Partial sums of the five MSMs over this rank's slice of every point range (contiguous split like
// multi_exp's chunks, multiexp.tcc:417-431; the last rank takes the remainder).
// h_r_fr / r_b1_out (optional, both or neither): r * (this call's B1 sum) is computed in B1's host tail, i.e. while the
// GPU is still busy with the L and H MSMs, instead of serially after the join (753 doublings on one core).
// spans: per query (0 A, 1 B1, 2 B2, 3 L, 4 H) the run [first, end) of `world` slices this call sums; an empty run
// leaves O in that slot (and, for H, skips the witness map).
static int prove_partials(b200_params *p, const void *h_input, size_t input_bytes, const int *spans, int world,
                          unsigned char *partials, b200_prove_timings *tm, const void *d_h_external = nullptr,
                          const unsigned char *h_r_fr = nullptr, unsigned char *r_b1_out = nullptr) {
  const size_t d = p->d, m = p->m;
  const size_t need = 96 * ((m + 1) + 3 * (d + 1) + 1);
  if (input_bytes != need) return set_error(-4, "input image has %zu bytes, expected %zu", input_bytes, need);
  B200_CHECK(check_spans(spans, world));
  const char *in = (const char *)h_input;
  double t0 = now_ms();
  // w first (it drives four of the five MSMs); ca/cb/cc follow asynchronously on the default stream right before
  // compute_H, under the MSMs that are already running (truly asynchronous when the image is in pinned memory)
  // (a rank of a sharded proof only needs the part of w its slices of A/B1/B2 (w[i]) and L (w[i+2]) read)
  {
    size_t wlo = m + 1, whi = 0;  // the union of the w ranges of the w-driven queries (L's point i reads w[i + 2])
    for (int qi = 0; qi < 4; qi++) {
      size_t lo, hi;
      query_slice(d, m, qi, spans[2 * qi], spans[2 * qi + 1], world, lo, hi);
      if (hi <= lo) continue;
      const size_t off = qi == 3 ? 2 : 0;
      if (lo + off < wlo) wlo = lo + off;
      if (hi + off > whi) whi = hi + off;
    }
    if (whi > wlo)
      B200_CUDA_CHECK(cudaMemcpy((char *)p->w.p + wlo * 96, in + wlo * 96, (whi - wlo) * 96, cudaMemcpyDefault));
  }
  double t1 = now_ms();
  const int curve = p->curve;
  const size_t g1a = affine_bytes(curve, 1), g2a = affine_bytes(curve, 2);
  const size_t g1p = proj_bytes(curve, 1), g2p = proj_bytes(curve, 2);
  if (use_precompute()) B200_CHECK(b200_params_precompute_queries(p, spans, world));
  struct Job { int group; const char *scalars; const char *points; size_t n; size_t stride; size_t outb; double *ms; };
  double ms[5] = {0, 0, 0, 0, 0};
  Job jobs[5] = {
      {1, (const char *)p->w.p, (const char *)p->q[0], m + 1, g1a, g1p, &ms[0]},        // A   main.cpp:227
      {1, (const char *)p->w.p, (const char *)p->q[1], m + 1, g1a, g1p, &ms[1]},        // B1  main.cpp:232
      {2, (const char *)p->w.p, (const char *)p->q[2], m + 1, g2a, g2p, &ms[2]},        // B2  main.cpp:237
      {1, d_h_external ? (const char *)d_h_external : (const char *)p->h.p, (const char *)p->q[4], d, g1a, g1p, &ms[3]},  // H   main.cpp:242
      {1, (const char *)p->w.p + 2 * 96, (const char *)p->q[3], m - 1, g1a, g1p, &ms[4]} // L   main.cpp:247 (w+2)
  };
  // The GPU half of each MSM runs here, back to back; the serial host halves (window combine, 753 doublings each)
  // run on worker threads while the next MSM occupies the GPU.
  // Issue order: B2 (G2, the longest), A, B1, L - all driven by w - each on its own stream; then compute_H on the
  // default stream, then the H MSM, which is fenced on the default stream. Measured alternatives (round 1,
  // tools/profile_shard.py timelines): compute_H first on a high-priority stream, and short kernels / accumulations
  // on separate high / low priority streams, both lengthen the proof (391 -> 407 / 411 ms): whatever runs beside an
  // accumulation kernel takes register-file space from it for longer than it saves.
  const size_t outoff[5] = {0, g1p, 2 * g1p, 2 * g1p + g2p, 3 * g1p + g2p};
  // B200_H_FIRST=1: the witness map and the H MSM are issued BEFORE the w-driven MSMs. When a rank's share of the points
  // is small (8 GPUs) compute_H - 21 short launches on the default stream - otherwise crawls behind the accumulation
  // blocks of four MSMs (13 ms alone, ~50 ms in their shadow) and the H MSM, which waits for it, ends the proof alone.
  static const int h_first_env = getenv("B200_H_FIRST") ? atoi(getenv("B200_H_FIRST")) : -1;
  const bool h_first = h_first_env >= 0 ? h_first_env != 0 : false;
  const int order_default[5] = {2, 0, 1, 4, 3}, order_h_first[5] = {3, 2, 0, 1, 4};
  int order[5];
  memcpy(order, h_first ? order_h_first : order_default, sizeof(order));
  int slot_of_job[5] = {1, 2, 0, 4, 3};  // workspace of A, B1, B2, H, L (slot 0 is the one the others share)
  const int jobq[5] = {0, 1, 2, 4, 3};   // job -> query index
  auto job_slice = [&](int j, size_t &lo, size_t &hi) {
    query_slice(d, m, jobq[j], spans[2 * jobq[j]], spans[2 * jobq[j] + 1], world, lo, hi);
  };
  // The preparation (digits, counting sort, task lists) of the w-driven MSMs is made once, in workspace 0, by B2 - or by
  // B1 when this call has no part in B2 (per-query sharding) - and reused by the others whose slice is the same range of w.
  int producer = 2;
  {
    size_t lo, hi;
    job_slice(2, lo, hi);
    if (hi <= lo) {
      producer = 1;
      slot_of_job[1] = 0;
      slot_of_job[2] = 2;
      for (int k = 0; k < 5; k++)  // the producer is issued before the MSMs that consume its preparation: B1 before A
        if (order[k] == 0 || order[k] == 1) order[k] ^= 1;
    }
  }
  // every tail reports its own status and message (it runs on another thread, whose thread-local error slot this thread
  // cannot see); a failed tail fails the proof
  struct TailResult { int rc = 0; std::string err; };
  std::vector<std::future<TailResult>> tails;
  int rc_all = 0;
  double t2 = t1;
  for (int jj = 0; jj < 5 && rc_all == 0; jj++) {
    const int j = order[jj];
    size_t lo, hi;
    job_slice(j, lo, hi);
    unsigned char *o = partials + outoff[j];
    if (hi <= lo) {  // no part in this MSM: O, so that any set of partial results sums to the proof
      B200_CHECK(group_zero(curve, j == 2 ? 2 : 1, o));
      if (j == 1 && r_b1_out && r_b1_out != o) B200_CHECK(group_zero(curve, 1, r_b1_out));
      continue;
    }
    if (j == 3 && !d_h_external) {  // H needs the witness map (unless the caller computed it elsewhere)
      double a = now_ms();
      const char *abc = in + (m + 1) * 96;
      B200_CUDA_CHECK(cudaMemcpyAsync(p->ca.p, abc, (d + 1) * 96, cudaMemcpyDefault, 0));
      B200_CUDA_CHECK(cudaMemcpyAsync(p->cb.p, abc + (d + 1) * 96, (d + 1) * 96, cudaMemcpyDefault, 0));
      B200_CUDA_CHECK(cudaMemcpyAsync(p->cc.p, abc + 2 * (d + 1) * 96, (d + 1) * 96, cudaMemcpyDefault, 0));
      rc_all = b200_compute_h(p->dom, p->ca.p, p->cb.p, p->cc.p, p->h.p);
      t2 = t1 + (now_ms() - a);
      if (rc_all) break;
    }
    const Job &J = jobs[j];
    double a = now_ms();
    MsmTail tail;
    msm_select_slot(slot_of_job[j]);
    // A and B1 (slots 1, 2) run over the same scalars and window plan as B2 (slot 0): they reuse its digits, counting
    // sort and task list
    // (not when this query's equal bases are merged: its scalars then differ from w)
    // L (slot 3) does too: its points are the same scalars shifted by 2 (main.cpp:247-250), so B2's entry list is
    // re-indexed for it instead of being rebuilt (MsmShare)
    const bool merges = use_precompute() && p->pre.dedup[j].merged > 0;
    MsmShare share;
    if (use_precompute() && !merges && j != producer && j != 3 && p->pre.plan[j].c == p->pre.plan[producer].c &&
        share_prep_enabled()) {
      size_t lo1, hi1;
      job_slice(producer, lo1, hi1);
      const bool producer_ok = hi1 > lo1 && !(p->pre.dedup[producer].merged > 0);
      if (producer_ok && (j == 0 || j == 1 || j == 2) && lo == lo1 && hi == hi1) share.slot = 0;
      // L's points [lo, hi) read w[lo + 2, hi + 2): inside the producer's range of w, and laid out as query_slice cuts it
      if (producer_ok && j == 4 && lo + 2 >= lo1 && hi + 2 <= hi1) {
        share.slot = 0;
        share.n_src = (uint32_t)(hi1 - lo1);
        share.shift = (uint32_t)(lo + 2 - lo1);
      }
    }
    if (use_precompute())
      rc_all = msm_table_dispatch_deferred(curve, J.group, J.scalars + lo * 96, p->pre.table[j].p, hi - lo,
                                           p->pre.plan[j], o, tail, share, &p->pre.dedup[j]);
    else
      rc_all = msm_dispatch_deferred(curve, J.group, J.scalars + lo * 96, J.points + lo * J.stride, hi - lo, o, tail);
    if (rc_all == 0) {
      const unsigned char *scale_by = (j == 1 && r_b1_out) ? h_r_fr : nullptr;  // r_b1_out may be o itself
      tails.push_back(std::async(std::launch::async, [tail, scale_by, o, r_b1_out, curve] {
        TailResult r;
        r.rc = tail(r.err);
        if (r.rc == 0 && scale_by && b200_g1_scale(curve, scale_by, o, r_b1_out) != 0) {
          r.rc = -1;
          r.err = "r * Bt1 failed";
        }
        return r;
      }));
    }
    *J.ms = now_ms() - a;
  }
  msm_select_slot(0);
  if (tl_all_issued_hook) {
    std::function<void()> *hook = tl_all_issued_hook;
    tl_all_issued_hook = nullptr;
    (*hook)();
  }
  double t_join = now_ms();
  std::string issue_err = rc_all ? last_error() : std::string();
  for (auto &f : tails) {
    TailResult r = f.get();
    if (r.rc && rc_all == 0) {
      rc_all = r.rc;
      issue_err = r.err;
    }
  }
  double join_ms = now_ms() - t_join;
  if (rc_all) return set_error(rc_all, "%s", issue_err.c_str());
  if (tm) {
    tm->h2d_ms = t1 - t0;
    tm->compute_h_ms = t2 - t1;
    tm->msm_a_ms = ms[0]; tm->msm_b1_ms = ms[1]; tm->msm_b2_ms = ms[2]; tm->msm_h_ms = ms[3]; tm->msm_l_ms = ms[4];
    tm->tail_ms = join_ms;
    tm->total_ms = now_ms() - t0;
  }
  return 0;
}

static size_t partial_size(int curve) { return 4 * proj_bytes(curve, 1) + proj_bytes(curve, 2); }

int b200_prove_partial(b200_params *p, const void *h_input, size_t input_bytes, int rank, int world, void *h_partials,
                       size_t *partial_bytes, b200_prove_timings *timings) {
  return b200_prove_partial_span(p, h_input, input_bytes, rank, rank + 1, world, h_partials, partial_bytes, timings);
}
int b200_prove_partial_span(b200_params *p, const void *h_input, size_t input_bytes, int rank, int rank_end, int world,
                            void *h_partials, size_t *partial_bytes, b200_prove_timings *timings) {
  return b200_prove_partial_ext(p, h_input, input_bytes, rank, rank_end, world, nullptr, h_partials, partial_bytes, timings);
}
int b200_prove_partial_ext(b200_params *p, const void *h_input, size_t input_bytes, int rank, int rank_end, int world,
                           const void *d_h_coefficients, void *h_partials, size_t *partial_bytes, b200_prove_timings *timings) {
  const int spans[10] = {rank, rank_end, rank, rank_end, rank, rank_end, rank, rank_end, rank, rank_end};
  return b200_prove_partial_queries(p, h_input, input_bytes, spans, world, 0, d_h_coefficients, h_partials, partial_bytes, timings);
}
int b200_prove_partial_queries(b200_params *p, const void *h_input, size_t input_bytes, const int *spans, int world,
                               int b1_scaled, const void *d_h_coefficients, void *h_partials, size_t *partial_bytes,
                               b200_prove_timings *timings) {
  B200_CHECK(require_device());
  unsigned char *part = (unsigned char *)h_partials;
  if (b1_scaled) {
    unsigned char r[96];
    if (input_bytes < 96) return set_error(-4, "input image has %zu bytes", input_bytes);
    B200_CUDA_CHECK(cudaMemcpy(r, (const unsigned char *)h_input + input_bytes - 96, 96, cudaMemcpyDefault));
    B200_CHECK(prove_partials(p, h_input, input_bytes, spans, world, part, timings, d_h_coefficients, r,
                              part + proj_bytes(p->curve, 1)));  // the B1 slot is scaled in place by its own tail
  } else {
    B200_CHECK(prove_partials(p, h_input, input_bytes, spans, world, part, timings, d_h_coefficients));
  }
  if (partial_bytes) *partial_bytes = partial_size(p->curve);
  return 0;
}
int b200_prove_partial_scaled(b200_params *p, const void *h_input, size_t input_bytes, int rank, int rank_end, int world,
                              const void *d_h_coefficients, void *h_partials, size_t *partial_bytes,
                              b200_prove_timings *timings) {
  const int spans[10] = {rank, rank_end, rank, rank_end, rank, rank_end, rank, rank_end, rank, rank_end};
  return b200_prove_partial_queries(p, h_input, input_bytes, spans, world, 1, d_h_coefficients, h_partials, partial_bytes, timings);
}

// C = Ht + Lt + r*Bt1 (main.cpp:253), then A | B | C in wire format (main.cpp:94-100)
// r_b1 (optional): r * Bt1 when the caller already has it (b200_prove computes it under the GPU work)
static int combine_partials(int curve, const void *h_partials_all, int world, const void *h_r_fr, const unsigned char *r_b1,
                            void *h_out, size_t *out_bytes) {
  if (curve != 0 && curve != 1) return set_error(-1, "bad curve %d", curve);
  const size_t g1p = proj_bytes(curve, 1), g2p = proj_bytes(curve, 2), ps = partial_size(curve);
  const size_t off[5] = {0, g1p, 2 * g1p, 2 * g1p + g2p, 3 * g1p + g2p};
  std::vector<unsigned char> sum(ps);
  const unsigned char *all = (const unsigned char *)h_partials_all;
  memcpy(sum.data(), all, ps);
  for (int r = 1; r < world; r++) {
    const unsigned char *pr = all + (size_t)r * ps;
    for (int j = 0; j < 5; j++) {
      if (j == 2) B200_CHECK(b200_g2_add(curve, sum.data() + off[j], pr + off[j], sum.data() + off[j]));
      else B200_CHECK(b200_g1_add(curve, sum.data() + off[j], pr + off[j], sum.data() + off[j]));
    }
  }
  std::vector<unsigned char> rb(g1p), c(g1p);
  if (r_b1) memcpy(rb.data(), r_b1, g1p);
  else if (!h_r_fr) memcpy(rb.data(), sum.data() + off[1], g1p);  // partials of b200_prove_partial_scaled
  else B200_CHECK(b200_g1_scale(curve, h_r_fr, sum.data() + off[1], rb.data()));
  B200_CHECK(b200_g1_add(curve, sum.data() + off[4], rb.data(), c.data()));
  B200_CHECK(b200_g1_add(curve, sum.data() + off[3], c.data(), c.data()));
  unsigned char *o = (unsigned char *)h_out;
  B200_CHECK(b200_g1_to_affine(curve, sum.data() + off[0], o));
  o += affine_bytes(curve, 1);
  B200_CHECK(b200_g2_to_affine(curve, sum.data() + off[2], o));
  o += affine_bytes(curve, 2);
  B200_CHECK(b200_g1_to_affine(curve, c.data(), o));
  o += affine_bytes(curve, 1);
  if (out_bytes) *out_bytes = (size_t)(o - (unsigned char *)h_out);
  return 0;
}
int b200_prove_combine(int curve, const void *h_partials_all, int world, const void *h_r_fr, void *h_out,
                       size_t *out_bytes) {
  return combine_partials(curve, h_partials_all, world, h_r_fr, nullptr, h_out, out_bytes);
}

int b200_prove(b200_params *p, const void *h_input, size_t input_bytes, void *h_out, size_t *out_bytes,
               b200_prove_timings *timings) {
  B200_CHECK(require_device());
  double t0 = now_ms();
  std::vector<unsigned char> part(partial_size(p->curve)), r_b1(proj_bytes(p->curve, 1));
  unsigned char r[96];  // the input image may live in host or device memory
  if (input_bytes < 96) return set_error(-4, "input image has %zu bytes", input_bytes);
  B200_CUDA_CHECK(cudaMemcpy(r, (const unsigned char *)h_input + input_bytes - 96, 96, cudaMemcpyDefault));
  const int whole[10] = {0, 1, 0, 1, 0, 1, 0, 1, 0, 1};
  B200_CHECK(prove_partials(p, h_input, input_bytes, whole, 1, part.data(), timings, nullptr, r, r_b1.data()));
  double t1 = now_ms();
  B200_CHECK(combine_partials(p->curve, part.data(), 1, r, r_b1.data(), h_out, out_bytes));
  if (timings) {
    timings->tail_ms += now_ms() - t1;
    timings->total_ms = now_ms() - t0;
  }
  return 0;
}

// One throw-away proof on pseudo-random inputs: loads every kernel's code (CUDA loads a kernel at its first launch),
// sizes the calling thread's five MSM workspaces, the pinned staging rings and the key's per-proof buffers for this key.
// Part of making a key ready (B::read_params calls it), like the base tables: a process that proves ONCE - the
// reference's driver - otherwise pays ~0.5 s of first-use costs inside its timed region.
int b200_params_warmup(b200_params *p) {
  B200_CHECK(require_device());
  const size_t n = (p->m + 1) + 3 * (p->d + 1) + 1;
  DevBuf img;
  B200_CHECK(img.alloc(n * 96));
  B200_CHECK(fr_fill_pseudo_random(img.p, n, 0xb200));
  B200_CUDA_CHECK(cudaDeviceSynchronize());
  std::vector<unsigned char> out(2 * affine_bytes(p->curve, 1) + affine_bytes(p->curve, 2));
  size_t nb = 0;
  return b200_prove(p, img.p, n * 96, out.data(), &nb, nullptr);
}

// ---- complete Groth16 proof terms: r1cs_gg_ppzksnark.tcc:457-470 / main.cpp:307-314
int b200_groth16_finalize(int curve, const void *h_proof, const void *h_r_fr, const void *h_s_fr, const void *h_extras,
                          void *h_out, size_t *out_bytes) {
  if (curve != 0 && curve != 1) return set_error(-1, "bad curve %d", curve);
  if (!h_proof || !h_r_fr || !h_s_fr || !h_extras || !h_out) return set_error(-1, "groth16_finalize: null argument");
  const size_t g1a = affine_bytes(curve, 1), g2a = affine_bytes(curve, 2), g1p = proj_bytes(curve, 1), g2p = proj_bytes(curve, 2);
  const unsigned char *pr = (const unsigned char *)h_proof, *ex = (const unsigned char *)h_extras;
  std::vector<unsigned char> A(g1p), B(g2p), C(g1p), alpha(g1p), beta1(g1p), delta1(g1p), beta2(g2p), delta2(g2p), t1(g1p), t2(g2p);
  B200_CHECK(b200_g1_from_affine(curve, pr, A.data()));
  B200_CHECK(b200_g2_from_affine(curve, pr + g1a, B.data()));
  B200_CHECK(b200_g1_from_affine(curve, pr + g1a + g2a, C.data()));
  B200_CHECK(b200_g1_from_affine(curve, ex, alpha.data()));
  B200_CHECK(b200_g1_from_affine(curve, ex + g1a, beta1.data()));
  B200_CHECK(b200_g1_from_affine(curve, ex + 2 * g1a, delta1.data()));
  B200_CHECK(b200_g2_from_affine(curve, ex + 3 * g1a, beta2.data()));
  B200_CHECK(b200_g2_from_affine(curve, ex + 3 * g1a + g2a, delta2.data()));
  // A' = alpha + A + r*delta
  B200_CHECK(b200_g1_scale(curve, h_r_fr, delta1.data(), t1.data()));
  B200_CHECK(b200_g1_add(curve, alpha.data(), A.data(), A.data()));
  B200_CHECK(b200_g1_add(curve, A.data(), t1.data(), A.data()));
  // B' = beta + B + s*delta
  B200_CHECK(b200_g2_scale(curve, h_s_fr, delta2.data(), t2.data()));
  B200_CHECK(b200_g2_add(curve, beta2.data(), B.data(), B.data()));
  B200_CHECK(b200_g2_add(curve, B.data(), t2.data(), B.data()));
  // C' = C + s*A' + r*beta_g1
  B200_CHECK(b200_g1_scale(curve, h_s_fr, A.data(), t1.data()));
  B200_CHECK(b200_g1_add(curve, C.data(), t1.data(), C.data()));
  B200_CHECK(b200_g1_scale(curve, h_r_fr, beta1.data(), t1.data()));
  B200_CHECK(b200_g1_add(curve, C.data(), t1.data(), C.data()));
  unsigned char *o = (unsigned char *)h_out;
  B200_CHECK(b200_g1_to_affine(curve, A.data(), o));
  B200_CHECK(b200_g2_to_affine(curve, B.data(), o + g1a));
  B200_CHECK(b200_g1_to_affine(curve, C.data(), o + g1a + g2a));
  if (out_bytes) *out_bytes = 2 * g1a + g2a;
  return 0;
}
int b200_prove_full(b200_params *p, const void *h_input, size_t input_bytes, const void *h_s_fr, const void *h_extras,
                    void *h_out, size_t *out_bytes) {
  std::vector<unsigned char> abc(2 * affine_bytes(p->curve, 1) + affine_bytes(p->curve, 2));
  size_t n = 0;
  B200_CHECK(b200_prove(p, h_input, input_bytes, abc.data(), &n, nullptr));
  unsigned char r[96];  // the input image may live in host or device memory
  B200_CUDA_CHECK(cudaMemcpy(r, (const unsigned char *)h_input + input_bytes - 96, 96, cudaMemcpyDefault));
  return b200_groth16_finalize(p->curve, abc.data(), r, h_s_fr, h_extras, h_out, out_bytes);
}

// ---- several proofs in flight at once -------------------------------------------------------------------------------
// A persistent host thread per extra job: the MSM workspaces and streams are thread_local, so a worker keeps its
// own set for the life of the process and its proof runs beside the caller's on the same GPU. The MNT6753 proof
// (2^15 constraints, latency-bound kernels) disappears under the MNT4753 one (2^20, multiplier-bound).
namespace {
class ProofWorker {
 public:
  ProofWorker() { std::thread([this] { loop(); }).detach(); }
  void submit(std::function<void()> job) {
    std::lock_guard<std::mutex> g(mu_);
    job_ = std::move(job);
    busy_ = true;
    cv_.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> g(mu_);
    cv_.wait(g, [this] { return !busy_; });
  }

 private:
  void loop() {
    msm_thread_high_priority(true);
    for (;;) {
      std::function<void()> job;
      {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [this] { return busy_ && job_; });
        job = std::move(job_);
        job_ = nullptr;
      }
      job();
      {
        std::lock_guard<std::mutex> g(mu_);
        busy_ = false;
      }
      cv_.notify_all();
    }
  }
  std::mutex mu_;
  std::condition_variable cv_;
  std::function<void()> job_;
  bool busy_ = false;
};

int run_proof_job(b200_proof_job *j) {
  if (!j->key || !j->h_input || !j->h_out) return set_error(-1, "proof job: null key, input or output");
  if (j->world > 1 && j->query_spans)
    return b200_prove_partial_queries(j->key, j->h_input, j->input_bytes, j->query_spans, j->world, j->b1_scaled,
                                      j->d_h_coefficients, j->h_out, &j->out_bytes, &j->timings);
  if (j->world > 1 && j->b1_scaled)
    return b200_prove_partial_scaled(j->key, j->h_input, j->input_bytes, j->rank, j->rank_end > j->rank ? j->rank_end : j->rank + 1,
                                     j->world, j->d_h_coefficients, j->h_out, &j->out_bytes, &j->timings);
  if (j->world > 1)
    return b200_prove_partial_ext(j->key, j->h_input, j->input_bytes, j->rank, j->rank_end > j->rank ? j->rank_end : j->rank + 1,
                                  j->world, j->d_h_coefficients, j->h_out, &j->out_bytes, &j->timings);
  return b200_prove(j->key, j->h_input, j->input_bytes, j->h_out, &j->out_bytes, &j->timings);
}
}  // namespace

int b200_prove_batch(b200_proof_job *jobs, int count) {
  B200_CHECK(require_device());
  // B200_BATCH_STAGGER=1 (off by default): job i+1 starts when job i has ISSUED all its MSMs instead of at time zero, so
  // that a small proof runs under the latency-bound end of a large one rather than beside its accumulations. Measured
  // on B200 (MNT4753 2^20 + MNT6753 2^15; alone 370 + 34 ms): started together 402 ms, staggered 414 ms - the host reaches
  // "all issued" only ~55 ms before the large proof ends, too late to hide a 34 ms latency-bound proof.
  static const bool stagger = getenv("B200_BATCH_STAGGER") && getenv("B200_BATCH_STAGGER")[0] == '1';
  struct ConcurrencyNote {  // several proofs in flight at once: the GPU is throughput-bound, see msm_use_coop
    explicit ConcurrencyNote(int k) { msm_note_concurrent_proofs(k); }
    ~ConcurrencyNote() { msm_note_concurrent_proofs(1); }
  } note(stagger ? 1 : count);
  constexpr int kMaxJobs = 8;
  if (!jobs || count < 1 || count > kMaxJobs) return set_error(-1, "prove_batch: count %d not in [1, %d]", count, kMaxJobs);
  for (int i = 0; i < count; i++)
    for (int k = i + 1; k < count; k++)
      if (jobs[i].key == jobs[k].key)
        return set_error(-1, "prove_batch: jobs %d and %d use the same key object (its per-proof buffers are not shared)", i, k);
  static std::mutex pool_mu;
  static ProofWorker *pool[kMaxJobs] = {nullptr};  // leaked on purpose: the threads outlive static destruction
  std::lock_guard<std::mutex> g(pool_mu);          // one batch at a time
  int dev = 0;
  B200_CUDA_CHECK(cudaGetDevice(&dev));
  std::string errs[kMaxJobs];
  // issued[i]: job i has issued all its MSMs (or ended before getting there)
  std::promise<void> issued[kMaxJobs];
  std::shared_future<void> issued_f[kMaxJobs];
  std::atomic<bool> issued_set[kMaxJobs];
  std::function<void()> hooks[kMaxJobs];
  for (int i = 0; i < count; i++) {
    issued_f[i] = issued[i].get_future().share();
    issued_set[i] = false;
    std::promise<void> *pr = &issued[i];
    std::atomic<bool> *flag = &issued_set[i];
    hooks[i] = [pr, flag] {
      if (!flag->exchange(true)) pr->set_value();
    };
  }
  for (int i = 1; i < count; i++) {
    if (!pool[i]) pool[i] = new ProofWorker();
    b200_proof_job *j = &jobs[i];
    std::string *err = &errs[i];
    std::function<void()> *hook = &hooks[i];
    std::shared_future<void> before = issued_f[i - 1];
    const bool wait_first = stagger;
    pool[i]->submit([j, err, dev, hook, before, wait_first] {
      if (wait_first) before.wait();
      cudaError_t e = cudaSetDevice(dev);
      tl_all_issued_hook = hook;
      j->status = e == cudaSuccess ? run_proof_job(j) : set_error(-2, "cudaSetDevice: %s", cudaGetErrorString(e));
      tl_all_issued_hook = nullptr;
      (*hook)();  // (a job that failed before issuing must not block its successor)
      if (j->status) *err = last_error();
    });
  }
  tl_all_issued_hook = &hooks[0];
  jobs[0].status = run_proof_job(&jobs[0]);
  tl_all_issued_hook = nullptr;
  hooks[0]();
  if (jobs[0].status) errs[0] = last_error();
  for (int i = 1; i < count; i++) pool[i]->wait();
  for (int i = 0; i < count; i++)
    if (jobs[i].status) return set_error(jobs[i].status, "proof job %d: %s", i, errs[i].c_str());
  return 0;
}

}  // extern "C"
