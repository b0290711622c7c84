// K1 — assembly kernels.  See assemble.cuh for the reference call sites replaced.
//
// Work decomposition (B200): one CTA owns ASM_ROWS_PER_CTA = 7 consecutive block rows, i.e.
// the 8 grid intervals touching them = 32 Gauss points = one lane per Gauss point.
//   phase 1  coalesced load of the sampled fields for those 32 Gauss points into shared memory
//   phase 2  spline values per Gauss point (warp 0)
//   phase 3  coefficient slots: warp w evaluates slots w, w+8, ... ; a slot is the sum of all
//            term factors sharing (matrix, row var, col var, d1, d2), times weight*dx
//   phase 4  per block row: half-warps expand (matrix, row var, col var) pairs into their 16
//            entries (spline1 * coef * spline2 summed over the 4 Gauss points), apply the
//            reference's per-contribution drop rule, combine the two diagonal-block
//            contributions, stage the 24 KB block row in shared memory and stream it out with
//            fully coalesced 16-byte stores.  No atomics on the matrix, no zero-fill pass.
// HBM traffic per block row: 2 x 12288 B written once, fields read once (+1/7 halo).

#include "assemble.cuh"

#include <algorithm>
#include <map>
#include <tuple>

namespace lgpu {

namespace {

enum Module : uint8_t {
  M_B, M_REG, M_FLOW, M_RES, M_HEAT, M_COND, M_VISC, M_HALLB, M_HALLA,
  N_REG, N_FLOW, N_RES, N_COND, N_VISC, N_HALLA, N_HALLB
};
enum Cond : uint8_t { C_NONE = 0, C_COMPR = 1, C_GRAV = 2, C_VHEAT = 4, C_INERTIA = 8, C_BFIELD = 16 };
enum Var : uint8_t { V_RHO, V_V1, V_V2, V_V3, V_TT, V_A1, V_A2, V_A3 };

struct TermDesc {
  uint8_t module, cond, v1, v2, d1, d2;
};

const TermDesc kTerms[] = {
#define T(M, C, V1, V2, D1, D2, E) {M, C, V_##V1, V_##V2, D1, D2},
#include "terms.def"
#undef T
};
constexpr int kNumTerms = sizeof(kTerms) / sizeof(kTerms[0]);

bool is_cubic(int var) { return var == V_V1 || var == V_A2 || var == V_A3; }
bool is_natural(int m) { return m >= N_REG; }
int matrix_of(int m) { return (m == M_B || m == M_HALLB || m == N_HALLB) ? 1 : 0; }

bool module_active(const lgpu_settings& s, int m) {
  const bool compr = !s.incompressible;
  switch (m) {
    case M_B: case M_REG: case N_REG: return true;
    case M_FLOW: case N_FLOW: return s.flow;
    case M_RES: case N_RES: return s.resistivity;
    case M_HEAT: return (s.cooling || s.heating) && compr;   // smod_heatloss_matrix.f08:12
    case M_COND: return s.conduction && compr;               // smod_conduction_matrix.f08:14
    case N_COND: return s.conduction;
    case M_VISC: case N_VISC: return s.viscosity;
    case M_HALLB: return s.hall;
    case M_HALLA: case N_HALLA: return s.hall && s.viscosity;  // smod_hall_matrix.f08:191 quirk
    case N_HALLB: return s.hall && s.electron_inertia;
  }
  return false;
}

bool cond_ok(const lgpu_settings& s, int c) {
  if ((c & C_COMPR) && s.incompressible) return false;
  if ((c & C_GRAV) && !s.gravity) return false;
  if ((c & C_VHEAT) && !(s.viscous_heating && !s.incompressible)) return false;
  if ((c & C_INERTIA) && !s.electron_inertia) return false;
  if ((c & C_BFIELD) && s.physics_type != 0) return false;
  return true;
}

// position of a variable inside the active state vector (src/settings/mod_settings.f08:69-86);
// -1: not in the state vector, every term naming it is skipped (mod_matrix_elements.f08:57-59).
// On the device every physics type is laid out in the 8-variable (16-wide) block structure: a
// present variable keeps its mhd slot, the slots of absent variables stay empty in A and B
// (the factorisation treats them as identity rows) and the C ABI translates indices and vectors
// to the reference's compact numbering (state_positions).
int var_position(int physics_type, int var) {
  static const int mhd[8] = {0, 1, 2, 3, 4, 5, 6, 7};
  static const int hd[8] = {0, 1, 2, 3, 4, -1, -1, -1};
  static const int hd1[8] = {0, 1, -1, -1, 2, -1, -1, -1};
  const int* tab = physics_type == 0 ? mhd : (physics_type == 1 ? hd : hd1);
  return tab[var];
}

}  // namespace

int state_positions(int physics_type, int slots[8]) {
  int n = 0;
  for (int v = 0; v < 8; ++v)
    if (var_position(physics_type, v) >= 0) slots[n++] = v;
  return n;
}

TermPlan build_term_plan(const lgpu_settings& s, bool natural) {
  TermPlan plan;
  std::map<std::tuple<int, int, int, int>, int> slot_of;   // (mat, p1, p2, dd) -> slot
  std::vector<std::vector<int32_t>> slot_terms;
  std::map<std::tuple<int, int, int>, int> item_of;        // (mat, p1, p2) -> item
  for (int id = 0; id < kNumTerms; ++id) {
    const TermDesc& t = kTerms[id];
    if (is_natural(t.module) != natural) continue;
    if (!module_active(s, t.module) || !cond_ok(s, t.cond)) continue;
    if (var_position(s.physics_type, t.v1) < 0 || var_position(s.physics_type, t.v2) < 0)
      continue;   // mod_matrix_elements.f08:57-59
    const int p1 = t.v1, p2 = t.v2;   // device slot = mhd position
    const int mat = matrix_of(t.module), dd = t.d1 * 2 + t.d2;
    auto skey = std::make_tuple(mat, p1, p2, dd);
    auto it = slot_of.find(skey);
    int slot;
    if (it == slot_of.end()) {
      slot = static_cast<int>(slot_terms.size());
      slot_of[skey] = slot;
      slot_terms.emplace_back();
      auto ikey = std::make_tuple(mat, p1, p2);
      auto jt = item_of.find(ikey);
      if (jt == item_of.end()) {
        PairItem item{};
        item.mat = mat; item.p1 = p1; item.p2 = p2;
        item.cls1 = is_cubic(t.v1) ? 2 : 0;
        item.cls2 = is_cubic(t.v2) ? 2 : 0;
        item.nslot = 0;
        item_of[ikey] = static_cast<int>(plan.items.size());
        plan.items.push_back(item);
        jt = item_of.find(ikey);
      }
      PairItem& item = plan.items[jt->second];
      item.slot[item.nslot] = slot;
      item.dd[item.nslot] = dd;
      ++item.nslot;
    } else {
      slot = it->second;
    }
    slot_terms[slot].push_back(id);
  }
  plan.slot_begin.push_back(0);
  for (auto& terms : slot_terms) {
    for (int id : terms) plan.term_ids.push_back(id);
    plan.slot_begin.push_back(static_cast<int32_t>(plan.term_ids.size()));
  }
  return plan;
}

std::vector<int32_t> essential_indices(const lgpu_settings& s, bool right_edge) {
  const int pt = s.physics_type;
  const int dsub = BLK;   // device layout: 16-wide blocks for every physics type
  auto idx = [&](std::initializer_list<int> vars, bool odd) {
    std::vector<int32_t> out;
    for (int v : vars) {
      if (var_position(pt, v) < 0) continue;
      int i = 2 * (v + 1);
      if (odd) --i;
      if (right_edge) i += dsub;
      out.push_back(i);
    }
    return out;
  };
  auto is_zero = [](double v) { return std::fabs(v) <= DP_LIMIT; };
  std::vector<int32_t> out;
  auto append = [&](const std::vector<int32_t>& v) { out.insert(out.end(), v.begin(), v.end()); };
  if (!right_edge) append(idx({V_RHO, V_V2, V_V3, V_TT, V_A1}, true));
  append(idx({V_V1}, true));
  if (s.boundary_type == 0) {
    append(idx({V_A2, V_A3}, true));
  } else {
    if (!is_zero(s.k2)) append(idx({V_A3}, true));
    if (!is_zero(s.k3)) append(idx({V_A2}, true));
  }
  if (s.perpendicular_conduction) append(idx({V_TT}, false));
  const bool noslip = s.viscosity && (right_edge || s.coaxial || s.geometry == 0);
  if (noslip) append(idx({V_V2, V_V3}, false));
  return out;
}

// ======================================================================== device side
namespace {

constexpr int SF_X = NFIELD;         // Gauss-point position
constexpr int SF_WDX = NFIELD + 1;   // quadrature weight * dx
constexpr int SF_ROWS = NFIELD + 2;

__device__ __forceinline__ cd tocd(double v) { return cd{v, 0.0}; }
__device__ __forceinline__ cd tocd(cd v) { return v; }
__device__ __forceinline__ double sq(double v) { return v * v; }

// ---- identifiers used by the FACTOR expressions of terms.def
#define FLD(i) sf[(i) * 32 + lane]
#define ic (cd{0.0, 1.0})
#define k2 (P.k2)
#define k3 (P.k3)
#define gamma_1 (P.gamma_1)
#define mu (P.mu)
#define efrac (P.efrac)
#define eps (P.geometry ? FLD(SF_X) : 1.0)
#define deps (P.geometry ? 1.0 : 0.0)
#define rho FLD(LGPU_F_RHO0)
#define drho FLD(LGPU_F_DRHO0)
#define T0 FLD(LGPU_F_T0)
#define dT0 FLD(LGPU_F_DT0)
#define ddT0 FLD(LGPU_F_DDT0)
#define B01 FLD(LGPU_F_B01)
#define B02 FLD(LGPU_F_B02)
#define dB02 FLD(LGPU_F_DB02)
#define ddB02 FLD(LGPU_F_DDB02)
#define B03 FLD(LGPU_F_B03)
#define dB03 FLD(LGPU_F_DB03)
#define ddB03 FLD(LGPU_F_DDB03)
#define v01 FLD(LGPU_F_V01)
#define dv01 FLD(LGPU_F_DV01)
#define ddv01 FLD(LGPU_F_DDV01)
#define v02 FLD(LGPU_F_V02)
#define dv02 FLD(LGPU_F_DV02)
#define ddv02 FLD(LGPU_F_DDV02)
#define v03 FLD(LGPU_F_V03)
#define dv03 FLD(LGPU_F_DV03)
#define ddv03 FLD(LGPU_F_DDV03)
#define g0 FLD(LGPU_F_G0)
#define eta FLD(LGPU_F_ETA)
#define detadT FLD(LGPU_F_DETADT)
#define deta (FLD(LGPU_F_DETADR) + (dT0 * detadT))
#define L0 FLD(LGPU_F_L0)
#define LT FLD(LGPU_F_DLDT)
#define Lrho FLD(LGPU_F_DLDRHO)
#define dkappa_para_dT FLD(LGPU_F_DTCPARADT)
#define kappa_perp FLD(LGPU_F_TCPERP)
#define dkappa_perp_drho FLD(LGPU_F_DTCPERPDRHO)
#define dkappa_perp_dT FLD(LGPU_F_DTCPERPDT)
#define dkappa_perp_dB2 FLD(LGPU_F_DTCPERPDB2)
#define Kp FLD(LGPU_F_TCPREFACTOR)
#define diffKp FLD(LGPU_F_DTCPREFACTORDR)
#define eta_H FLD(LGPU_F_HALLFACTOR)
#define eta_e FLD(LGPU_F_INERTIAFACTOR)
// derived operators (same definitions as the reference procedures)
#define drB02 (deps * B02 + eps * dB02)
#define Fop_plus (k2 * B02 / eps + k3 * B03)
#define Gop_plus (k3 * B02 + k2 * B03 / eps)
#define Gop_min (k3 * B02 - k2 * B03 / eps)
#define WVop (sq(k2) / eps + eps * sq(k3))
#define drv01 (deps * v01 + eps * dv01)
#define drv02 (deps * v02 + eps * dv02)
#define Vop (k2 * v02 / eps + k3 * v03)
#define Rop_pos (deps * eta / eps + deta)
#define Rop_neg (deps * eta / eps - deta)
#define B0sq (sq(B01) + sq(B02) + sq(B03))
#define Kp_plus (Kp + dkappa_perp_dB2)
#define Kp_plusplus (dkappa_perp_dB2 - (sq(B01) * Kp_plus / B0sq))
#define dFop_plus ((k2 / eps) * (dB02 - deps * B02 / eps) + k3 * dB03)
#define Fop_B01 (deps * ic * B01 / eps + Fop_plus)

enum { kCaseBase = __COUNTER__ + 1 };

// Closed-form factor of term `id` at the Gauss point held by `lane` (fields in shared memory).
__device__ __noinline__ cd term_factor(int id, const double* __restrict__ sf, int lane,
                                       const AsmParams& P) {
  switch (id) {
#define T(M, C, V1, V2, D1, D2, E) \
  case (__COUNTER__ - kCaseBase):  \
    return tocd(E);
#include "terms.def"
#undef T
  }
  return cd{0.0, 0.0};
}

#undef FLD
#undef ic
#undef k2
#undef k3
#undef gamma_1
#undef mu
#undef efrac
#undef eps
#undef deps
#undef rho
#undef drho
#undef T0
#undef dT0
#undef ddT0
#undef B01
#undef B02
#undef dB02
#undef ddB02
#undef B03
#undef dB03
#undef ddB03
#undef v01
#undef dv01
#undef ddv01
#undef v02
#undef dv02
#undef ddv02
#undef v03
#undef dv03
#undef ddv03
#undef g0
#undef eta
#undef detadT
#undef deta
#undef L0
#undef LT
#undef Lrho
#undef dkappa_para_dT
#undef kappa_perp
#undef dkappa_perp_drho
#undef dkappa_perp_dT
#undef dkappa_perp_dB2
#undef Kp
#undef diffKp
#undef eta_H
#undef eta_e
#undef drB02
#undef Fop_plus
#undef Gop_plus
#undef Gop_min
#undef WVop
#undef drv01
#undef drv02
#undef Vop
#undef Rop_pos
#undef Rop_neg
#undef B0sq
#undef Kp_plus
#undef Kp_plusplus
#undef dFop_plus
#undef Fop_B01

// spl[(cls*4 + entry)*32 + lane]; cls: 0 h_quad, 1 dh_quad, 2 h_cubic, 3 dh_cubic
// (src/mod_spline_functions.f08:23-99, same expression order)
__device__ void eval_splines(double r, double lo, double hi, double* spl, int lane) {
  const double h = hi - lo, h2 = h * h, h3 = h2 * h;
  double v[16];
  v[0] = 4.0 * (r - lo) * (hi - r) / h2;
  v[1] = 0.0;
  v[2] = (2.0 * r - hi - lo) * (r - lo) / h2;
  v[3] = (2.0 * r - hi - lo) * (r - hi) / h2;
  v[4] = 4.0 * (-2.0 * r + hi + lo) / h2;
  v[5] = 0.0;
  v[6] = (4.0 * r - hi - 3.0 * lo) / h2;
  v[7] = (4.0 * r - lo - 3.0 * hi) / h2;
  const double a = (r - lo) / h, b = (hi - r) / h;
  v[8] = 3.0 * a * a - 2.0 * a * a * a;
  v[9] = 3.0 * b * b - 2.0 * b * b * b;
  v[10] = (r - hi) * a * a;
  v[11] = (r - lo) * b * b;
  v[12] = 6.0 * (r - lo) / h2 - 6.0 * (r - lo) * (r - lo) / h3;
  v[13] = -6.0 * (hi - r) / h2 + 6.0 * (hi - r) * (hi - r) / h3;
  v[14] = (2.0 * (r - hi) * (r - lo) + (r - lo) * (r - lo)) / h2;
  v[15] = (2.0 * (r - lo) * (r - hi) + (r - hi) * (r - hi)) / h2;
#pragma unroll
  for (int i = 0; i < 16; ++i) spl[i * 32 + lane] = v[i];
}

__device__ __forceinline__ bool dropped(cd v) {   // mod_check_values.f08:143-192 (is_zero)
  return fabs(v.x) <= DP_LIMIT && fabs(v.y) <= DP_LIMIT;
}

// Sum over Gauss points [g0, g0 + ng) and the item's slots of spline1 * coef * spline2.
__device__ __forceinline__ cd item_entry(const PairItem& it, const double* spl, const cd* coef,
                                         int e1, int e2, int gfirst, int ng) {
  cd val{0.0, 0.0};
  for (int si = 0; si < it.nslot; ++si) {
    const int c1 = it.cls1 + (it.dd[si] >> 1), c2 = it.cls2 + (it.dd[si] & 1);
    const double* s1 = spl + (c1 * 4 + e1) * 32;
    const double* s2 = spl + (c2 * 4 + e2) * 32;
    const cd* cf = coef + it.slot[si] * 32;
    for (int g = gfirst; g < gfirst + ng; ++g) {
      const cd c = cf[g];
      const double a = s1[g], b = s2[g];
      val.x += (a * c.x) * b;
      val.y += (a * c.y) * b;
    }
  }
  return val;
}

__global__ void __launch_bounds__(256)
assemble_kernel(AsmParams P, DevicePlan plan, FieldPtrs fields, const double* __restrict__ grid,
                const double* __restrict__ gauss, cd* __restrict__ A, cd* __restrict__ B,
                uint32_t* __restrict__ masks) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd* tile = reinterpret_cast<cd*>(smem_raw);                      // [2][3][256]
  cd* coef = tile + 2 * 3 * BLK2;                                  // [nslots][32]
  double* sf = reinterpret_cast<double*>(coef + plan.nslots * 32); // [SF_ROWS][32]
  double* spl = sf + SF_ROWS * 32;                                 // [16][32]
  uint32_t* smask = reinterpret_cast<uint32_t*>(spl + 16 * 32);    // [MASK_WORDS]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = P.gridpts;
  const int b0 = blockIdx.x * ASM_ROWS_PER_CTA;   // first block row of this CTA
  const int elem = b0 - 1 + (lane >> 2);          // grid interval of this lane's Gauss point
  const bool evalid = elem >= 0 && elem <= G - 2;
  const int gp = 4 * elem + (lane & 3);

  // phase 1: sampled fields -> shared memory (coalesced: consecutive lanes, consecutive points)
  for (int f = warp; f < NFIELD; f += 8) {
    const double* src = fields.f[f];
    sf[f * 32 + lane] = (evalid && src) ? __ldg(src + gp) : 0.0;
  }
  if (warp == 0) {
    // phase 2: position, weight*dx and the 16 spline values of this Gauss point
    double x = 1.0, wdx = 0.0, lo = 0.0, hi = 1.0;
    if (evalid) {
      x = __ldg(gauss + gp);
      lo = __ldg(grid + elem);
      hi = __ldg(grid + elem + 1);
      wdx = P.weights[lane & 3] * (hi - lo);
    }
    sf[SF_X * 32 + lane] = x;
    sf[SF_WDX * 32 + lane] = wdx;
    eval_splines(x, lo, hi, spl, lane);
  }
  __syncthreads();

  // phase 3: coefficient slots (uniform control flow inside a warp: lanes = Gauss points)
  for (int s = warp; s < plan.nslots; s += 8) {
    cd acc{0.0, 0.0};
    const int tb = __ldg(plan.slot_begin + s), te = __ldg(plan.slot_begin + s + 1);
    for (int t = tb; t < te; ++t) acc += term_factor(__ldg(plan.term_ids + t), sf, lane, P);
    coef[s * 32 + lane] = acc * sf[SF_WDX * 32 + lane];
  }
  __syncthreads();

  // phase 4: expand pairs into entries, one block row at a time
  const int hw = tid >> 4;            // half-warp id 0..15
  const int hl = tid & 15;            // lane inside the half-warp
  const int eb = hl >> 3;             // 0: element b-1 (bottom rows), 1: element b (top rows)
  const int ra = (hl >> 2) & 1;       // row parity inside the variable's 2x2
  const int cq = (hl >> 1) & 1;       // 0: left columns, 1: right columns of the quadblock
  const int cb = hl & 1;              // column parity
  const int e1 = eb ? 2 * ra + 1 : 2 * ra;   // spline entry (0-based): top rows use (2,4), bottom (1,3)
  const int e2 = cq ? 2 * cb : 2 * cb + 1;
  const int contrib = eb * 2 + cq;    // 0 sub, 1 diag (from b-1), 2 diag (from b), 3 super
  const int nitems_pad = (plan.nitems + 15) & ~15;

  for (int r = 0; r < ASM_ROWS_PER_CTA; ++r) {
    const int b = b0 + r;
    if (b >= G) break;   // uniform
    for (int i = tid; i < 2 * 3 * BLK2; i += 256) tile[i] = cd{0.0, 0.0};
    if (tid < MASK_WORDS) smask[tid] = 0u;
    __syncthreads();
    const int le = r + eb;                         // local element index 0..7
    const int e_glob = b0 - 1 + le;
    const bool ev = e_glob >= 0 && e_glob <= G - 2;
    for (int i = hw; i < nitems_pad; i += 16) {
      const bool active = i < plan.nitems;
      cd val{0.0, 0.0};
      int mat = 0, idx = 0;
      if (active) {
        const PairItem it = plan.items[i];
        mat = it.mat;
        idx = (2 * it.p2 + cb) * BLK + 2 * it.p1 + ra;   // column-major inside the block
        if (ev) val = item_entry(it, spl, coef, e1, e2, le * 4, 4);
      }
      const bool keep = active && !dropped(val);
      if (!keep) val = cd{0.0, 0.0};
      // diag block = contribution of element b-1 (lanes contrib 1) + element b (contrib 2, +6 lanes)
      const double ox = __shfl_down_sync(0xffffffffu, val.x, 6, 16);
      const double oy = __shfl_down_sync(0xffffffffu, val.y, 6, 16);
      if (active) {
        cd* dst = tile + mat * 3 * BLK2;
        if (contrib == 0) dst[idx] = val;
        else if (contrib == 1) dst[BLK2 + idx] = cd{val.x + ox, val.y + oy};
        else if (contrib == 3) dst[2 * BLK2 + idx] = val;
        if (keep) atomicOr(&smask[mat * 32 + contrib * 8 + (idx >> 5)], 1u << (idx & 31));
      }
    }
    __syncthreads();
    cd* dA = A + static_cast<size_t>(b) * 3 * BLK2;
    cd* dB = B + static_cast<size_t>(b) * 3 * BLK2;
    for (int i = tid; i < 3 * BLK2; i += 256) {
      dA[i] = tile[i];
      dB[i] = tile[3 * BLK2 + i];
    }
    if (tid < MASK_WORDS) masks[static_cast<size_t>(b) * MASK_WORDS + tid] = smask[tid];
    __syncthreads();
  }
}

// Locate global entry (gr, gc) in the block-tridiagonal layout.
__device__ __forceinline__ size_t entry_offset(int gr, int gc, int* tile_out) {
  const int b = gr / BLK, t = gc / BLK - b + 1;
  *tile_out = t;
  return (static_cast<size_t>(b) * 3 + t) * BLK2 + (gc % BLK) * BLK + gr % BLK;
}

// Natural + essential boundary conditions (src/boundaries/*): one CTA, both edges in turn.
__global__ void __launch_bounds__(256)
boundary_kernel(AsmParams P, DevicePlan plan, FieldPtrs fields, const double* __restrict__ grid,
                const double* __restrict__ gauss, cd* A, cd* B, uint32_t* masks,
                uint32_t* natmasks, const int32_t* ess_left, int n_left, const int32_t* ess_right,
                int n_right) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd* coef = reinterpret_cast<cd*>(smem_raw);                       // [nslots][32] (lane 0 used)
  double* sf = reinterpret_cast<double*>(coef + plan.nslots * 32);  // [SF_ROWS][32]
  double* spl = sf + SF_ROWS * 32;
  const int tid = threadIdx.x;
  const int G = P.gridpts;
  const int dimq = 2 * BLK;

  for (int edge = 0; edge < 2; ++edge) {
    const int gp = edge == 0 ? 0 : 4 * (G - 1) - 1;       // first / last Gaussian point
    const int e0 = edge == 0 ? 0 : G - 2;                 // first / last interval
    const double weight = edge == 0 ? -1.0 : 1.0;         // Bounds[x1] - Bounds[x0]
    const int shift = edge == 0 ? 0 : (G - 2) * BLK;      // global index of the quadblock corner
    if (tid < NFIELD) sf[tid * 32] = fields.f[tid] ? fields.f[tid][gp] : 0.0;
    if (tid == 0) {
      sf[SF_X * 32] = gauss[gp];
      const double lo = grid[e0], hi = grid[e0 + 1];
      eval_splines(edge == 0 ? lo : hi, lo, hi, spl, 0);   // basis functions at the actual edge
    }
    __syncthreads();
    for (int s = tid; s < plan.nslots; s += 256) {
      cd acc{0.0, 0.0};
      for (int t = plan.slot_begin[s]; t < plan.slot_begin[s + 1]; ++t)
        acc += term_factor(plan.term_ids[t], sf, 0, P);
      coef[s * 32] = acc * weight;
    }
    __syncthreads();
    // natural quadblock entries: 16 per (matrix, row var, col var) pair
    const int hl = tid & 15;
    const int qr = hl >> 3, ra = (hl >> 2) & 1, qc = (hl >> 1) & 1, cb = hl & 1;
    const int e1 = qr ? 2 * ra : 2 * ra + 1;   // top rows use entries (2,4), bottom rows (1,3)
    const int e2 = qc ? 2 * cb : 2 * cb + 1;
    for (int i = tid >> 4; i < plan.nitems; i += 16) {
      const PairItem it = plan.items[i];
      const cd val = item_entry(it, spl, coef, e1, e2, 0, 1);
      if (dropped(val)) continue;
      const int row = 2 * it.p1 + ra, col = 2 * it.p2 + cb;
      const int gr = shift + qr * BLK + row, gc = shift + qc * BLK + col;
      int t;
      const size_t off = entry_offset(gr, gc, &t);
      cd* M = it.mat ? B : A;
      M[off] += val;
      const int idx = col * BLK + row;
      atomicOr(&natmasks[((edge * 2 + it.mat) * 4 + qr * 2 + qc) * 8 + (idx >> 5)], 1u << (idx & 31));
    }
    __syncthreads();
    // essential conditions: wipe row + column inside the edge quadblock, then B_ii = 1, A_ii = 0
    const int32_t* ess = edge == 0 ? ess_left : ess_right;
    const int ness = edge == 0 ? n_left : n_right;
    for (int w = tid; w < ness * dimq * 2 * 2; w += 256) {
      const int mat = w & 1, dir = (w >> 1) & 1, k = (w >> 2) % dimq, n = (w >> 2) / dimq;
      const int g = shift + ess[n] - 1;
      const int gr = dir ? shift + k : g, gc = dir ? g : shift + k;
      int t;
      const size_t off = entry_offset(gr, gc, &t);
      (mat ? B : A)[off] = cd{0.0, 0.0};
      // clear every structural bit of this entry
      const int b = gr / BLK, idx = (gc % BLK) * BLK + gr % BLK;
      const uint32_t bit = ~(1u << (idx & 31));
      uint32_t* mw = masks + static_cast<size_t>(b) * MASK_WORDS + mat * 32 + (idx >> 5);
      if (t == 0) atomicAnd(mw + 0 * 8, bit);
      if (t == 1) { atomicAnd(mw + 1 * 8, bit); atomicAnd(mw + 2 * 8, bit); }
      if (t == 2) atomicAnd(mw + 3 * 8, bit);
      const int qr2 = (gr - shift) / BLK, qc2 = (gc - shift) / BLK;
      atomicAnd(&natmasks[((edge * 2 + mat) * 4 + qr2 * 2 + qc2) * 8 + (idx >> 5)], bit);
    }
    __syncthreads();
    for (int n = tid; n < ness; n += 256) {
      const int g = shift + ess[n] - 1;
      int t;
      const size_t off = entry_offset(g, g, &t);
      B[off] = cd{1.0, 0.0};
      // essential-diagonal marker: natmasks tail [2 edges][32 bits], B only (A's 0 is dropped)
      atomicOr(&natmasks[2 * 2 * 4 * 8 + edge], 1u << (ess[n] - 1));
    }
    __syncthreads();
  }
}

}  // namespace

size_t assemble_smem_bytes(int nslots) {
  return sizeof(cd) * (2 * 3 * BLK2 + static_cast<size_t>(nslots) * 32) +
         sizeof(double) * (SF_ROWS * 32 + 16 * 32) + sizeof(uint32_t) * MASK_WORDS;
}

void launch_assemble(const AsmParams& p, const DevicePlan& plan, const FieldPtrs& fields,
                     const double* grid, const double* gauss_grid, cd* A, cd* B, uint32_t* masks,
                     cudaStream_t stream) {
  const size_t smem = assemble_smem_bytes(plan.nslots);
  CUDA_CHECK(cudaFuncSetAttribute(assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
  const int ctas = (p.gridpts + ASM_ROWS_PER_CTA - 1) / ASM_ROWS_PER_CTA;
  assemble_kernel<<<ctas, 256, smem, stream>>>(p, plan, fields, grid, gauss_grid, A, B, masks);
  CUDA_CHECK(cudaGetLastError());
}

void launch_boundaries(const AsmParams& p, const DevicePlan& natplan, const FieldPtrs& fields,
                       const double* grid, const double* gauss_grid, cd* A, cd* B, uint32_t* masks,
                       uint32_t* natmasks, const int32_t* ess_left, int n_left,
                       const int32_t* ess_right, int n_right, cudaStream_t stream) {
  const size_t smem = sizeof(cd) * static_cast<size_t>(natplan.nslots) * 32 +
                      sizeof(double) * (SF_ROWS * 32 + 16 * 32);
  boundary_kernel<<<1, 256, smem, stream>>>(p, natplan, fields, grid, gauss_grid, A, B, masks,
                                            natmasks, ess_left, n_left, ess_right, n_right);
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace lgpu
