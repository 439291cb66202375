// Host-side construction and verification of the atomic-free scatter plans
// (see wfx_plan.h).  Pure C++: runs without a GPU, so the plans are unit-tested on CPU.
#include "wfx_plan.h"
#include "wfx_internal.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>

namespace wfx
{
namespace
{
inline int lowest_zero_bit(uint64_t m)
{
  if (~m == 0) return -1;
  return __builtin_ctzll(~m);
}

// ascending lattice position of 1-D dof a in the [0, 1, interior] ordering
inline int ascpos(int a, int P) { return a == 0 ? 0 : (a == 1 ? P : a - 1); }

// Strides (Sx, Sy) of the in-brick placement X*Sx + Y*Sy + Z: injective on the (P*be+1)^3 lattice
// and, if possible, such that the lanes (i,j) of a cell hit distinct shared-memory banks when they
// access one k-plane.  word_bytes = 8: 16 banks of 8 bytes per half-warp; 4: 32 banks per warp.
struct Strides
{
  bool ok = false, tuned = false;
  int Sx = 0, Sy = 0;
};
Strides find_strides(int P, const BrickShape& brick, int word_bytes, bool tuned)
{
  const int n = P + 1;
  const int Ex = P * brick.e[0] + 1, Ey = P * brick.e[1] + 1, Ez = P * brick.e[2] + 1; // lattice extents
  const int banks = word_bytes == 8 ? 16 : 32;
  Strides best;
  // Degree 4, 64-bit words, cubic bricks: the kernel has a layout in which ALL of a cell's
  // shared-memory accesses (three roles, dof arrays and tiles) are conflict-free; it needs Sx = 5,
  // Sy = 2 (mod 16), see tools/bank_layout_search.py.  Costs 7 % padding of the dof arrays.
  if (tuned && P == 4 && word_bytes == 8 && brick.cubic())
  {
    best.Sy = Ez;
    while (best.Sy % 16 != 2) ++best.Sy;
    best.Sx = (Ey - 1) * best.Sy + Ez;
    while (best.Sx % 16 != 5) ++best.Sx;
    best.ok = best.tuned = true;
    return best;
  }
  const int dense = Ex * Ey * Ez;
  int best_conf = 1 << 30, best_cap = 1 << 30;
  for (int Sy = Ez; Sy < Ez + 32; ++Sy)
    for (int Sx = (Ey - 1) * Sy + Ez; Sx < (Ey - 1) * Sy + Ez + 32; ++Sx)
    {
      int conf = 0;
      for (int g0 = 0; g0 < n * n; g0 += banks)
      {
        int cnt[32] = {0};
        for (int lane = g0; lane < std::min(n * n, g0 + banks); ++lane)
        {
          const int bk = (ascpos(lane / n, P) * Sx + ascpos(lane % n, P) * Sy) % banks;
          conf += cnt[bk]++;
        }
      }
      const int cap = (Ex - 1) * Sx + (Ey - 1) * Sy + Ez;
      if (cap > dense + dense / 25) continue; // at most 4 % padding
      if (conf < best_conf || (conf == best_conf && cap < best_cap))
      {
        best_conf = conf;
        best_cap = cap;
        best.Sx = Sx;
        best.Sy = Sy;
        best.ok = true;
      }
    }
  return best;
}
} // namespace

void parallel_for(int64_t n, const std::function<void(int64_t, int64_t)>& fn, int64_t min_parallel)
{
  int nt = (int)std::thread::hardware_concurrency();
  if (const char* e = std::getenv("WFX_HOST_THREADS")) nt = std::atoi(e);
  nt = std::max(1, std::min(nt, 32));
  if (n < min_parallel || nt == 1)
  {
    fn(0, n);
    return;
  }
  std::vector<std::thread> th;
  std::vector<std::exception_ptr> err((size_t)nt);
  const int64_t chunk = (n + nt - 1) / nt;
  for (int t = 0; t < nt; ++t)
  {
    const int64_t b = t * chunk, e = std::min(n, b + chunk);
    if (b >= e) break;
    th.emplace_back([&, t, b, e] {
      try { fn(b, e); }
      catch (...) { err[t] = std::current_exception(); }
    });
  }
  for (auto& x : th) x.join();
  for (auto& e : err)
    if (e) std::rethrow_exception(e);
}

bool structured_cell_coords(int64_t ncells, int64_t npts, const int32_t* xdofs, std::vector<int32_t>& ijk)
{
  if (ncells <= 0 || npts <= 0) return false;
  // vertex -> cells (CSR)
  std::vector<int32_t> voff((size_t)npts + 1, 0);
  for (int64_t q = 0; q < ncells * 8; ++q) voff[xdofs[q] + 1]++;
  for (int64_t v = 0; v < npts; ++v) voff[v + 1] += voff[v];
  std::vector<int32_t> vcell((size_t)ncells * 8);
  {
    std::vector<int32_t> pos(voff.begin(), voff.end() - 1);
    for (int64_t c = 0; c < ncells; ++c)
      for (int v = 0; v < 8; ++v) vcell[pos[xdofs[8 * c + v]]++] = (int32_t)c;
  }
  // face neighbours: face (axis a, side s) = the 4 vertices whose bit a equals s
  std::vector<int32_t> nbr((size_t)ncells * 6, -1);
  std::vector<uint8_t> bad(1, 0);
  parallel_for(ncells, [&](int64_t c0, int64_t c1) {
    for (int64_t c = c0; c < c1 && !bad[0]; ++c)
    {
      const int32_t* xc = xdofs + 8 * c;
      for (int a = 0; a < 3; ++a)
        for (int s = 0; s < 2; ++s)
        {
          int fv[4], nf = 0;
          for (int v = 0; v < 8; ++v)
            if (((v >> a) & 1) == s) fv[nf++] = v;
          const int32_t v0 = xc[fv[0]];
          int32_t found = -1;
          for (int32_t p = voff[v0]; p < voff[v0 + 1] && found < 0; ++p)
          {
            const int32_t n = vcell[p];
            if (n == c) continue;
            const int32_t* xn = xdofs + 8 * (int64_t)n;
            int hits = 0;
            for (int q = 0; q < 4; ++q)
              for (int w = 0; w < 8; ++w)
                if (xn[w] == xc[fv[q]]) { ++hits; break; }
            if (hits == 4) found = n;
          }
          if (found >= 0)
          {
            // same orientation: my vertex v of the face is the neighbour's vertex v with bit a flipped
            const int32_t* xn = xdofs + 8 * (int64_t)found;
            for (int q = 0; q < 4; ++q)
              if (xn[fv[q] ^ (1 << a)] != xc[fv[q]]) bad[0] = 1;
          }
          nbr[6 * c + 2 * a + s] = found;
        }
    }
  });
  if (bad[0]) return false;
  // breadth-first coordinates
  std::vector<int32_t> co((size_t)ncells * 3, 0);
  std::vector<uint8_t> seen((size_t)ncells, 0);
  std::vector<int32_t> queue;
  queue.reserve((size_t)ncells);
  queue.push_back(0);
  seen[0] = 1;
  for (size_t h = 0; h < queue.size(); ++h)
  {
    const int32_t c = queue[h];
    for (int a = 0; a < 3; ++a)
      for (int s = 0; s < 2; ++s)
      {
        const int32_t n = nbr[6 * (int64_t)c + 2 * a + s];
        if (n < 0) continue;
        int32_t want[3] = {co[3 * (int64_t)c], co[3 * (int64_t)c + 1], co[3 * (int64_t)c + 2]};
        want[a] += s ? 1 : -1;
        if (!seen[n])
        {
          seen[n] = 1;
          for (int q = 0; q < 3; ++q) co[3 * (int64_t)n + q] = want[q];
          queue.push_back(n);
        }
        else
          for (int q = 0; q < 3; ++q)
            if (co[3 * (int64_t)n + q] != want[q]) return false; // not a lattice (e.g. an O-grid)
      }
  }
  if ((int64_t)queue.size() != ncells) return false; // more than one block
  int32_t lo[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, hi[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
  for (int64_t c = 0; c < ncells; ++c)
    for (int q = 0; q < 3; ++q)
    {
      lo[q] = std::min(lo[q], co[3 * c + q]);
      hi[q] = std::max(hi[q], co[3 * c + q]);
    }
  for (int q = 0; q < 3; ++q)
    if ((int64_t)hi[q] - lo[q] > 0xFFFF) return false;
  // two cells must not share a coordinate triple (a periodic mesh would wrap around)
  const int64_t e1 = hi[1] - lo[1] + 1, e2 = hi[2] - lo[2] + 1, e0 = hi[0] - lo[0] + 1;
  if (e0 * e1 * e2 > 8 * ncells + 64) return false; // far from a block: do not allocate the check
  std::vector<uint8_t> occ((size_t)(e0 * e1 * e2), 0);
  for (int64_t c = 0; c < ncells; ++c)
  {
    const int64_t key = ((int64_t)(co[3 * c] - lo[0]) * e1 + (co[3 * c + 1] - lo[1])) * e2 + (co[3 * c + 2] - lo[2]);
    if (occ[key]) return false;
    occ[key] = 1;
  }
  for (int64_t c = 0; c < ncells; ++c)
    for (int q = 0; q < 3; ++q) co[3 * c + q] -= lo[q];
  ijk.swap(co);
  return true;
}

void build_cell_colour_plan(int nd, int64_t ncells, int64_t ndofs, const int32_t* tdm,
                            CellColourPlan& plan)
{
  std::vector<uint64_t> dofmask((size_t)ndofs, 0);
  std::vector<uint8_t> colour((size_t)ncells);
  int ncol = 0;
  for (int64_t c = 0; c < ncells; ++c)
  {
    const int32_t* d = tdm + c * nd;
    uint64_t m = 0;
    for (int t = 0; t < nd; ++t) m |= dofmask[d[t]];
    const int col = lowest_zero_bit(m);
    if (col < 0) fail("cell colouring needs more than 64 colours");
    const uint64_t bit = 1ull << col;
    for (int t = 0; t < nd; ++t) dofmask[d[t]] |= bit;
    colour[c] = (uint8_t)col;
    ncol = std::max(ncol, col + 1);
  }
  plan.ncolours = ncol;
  plan.colour_off.assign(ncol + 1, 0);
  for (int64_t c = 0; c < ncells; ++c) plan.colour_off[colour[c] + 1]++;
  for (int k = 0; k < ncol; ++k) plan.colour_off[k + 1] += plan.colour_off[k];
  plan.cells.resize((size_t)ncells);
  std::vector<int32_t> pos(plan.colour_off.begin(), plan.colour_off.end() - 1);
  for (int64_t c = 0; c < ncells; ++c) plan.cells[pos[colour[c]]++] = (int32_t)c;
}

void verify_cell_colour_plan(const CellColourPlan& plan, int nd, int64_t ncells, int64_t ndofs,
                             const int32_t* tdm)
{
  if ((int64_t)plan.cells.size() != ncells) fail("colour plan: wrong cell count");
  std::vector<int32_t> stamp((size_t)ndofs, -1);
  std::vector<uint8_t> seen((size_t)ncells, 0);
  for (int k = 0; k < plan.ncolours; ++k)
    for (int32_t p = plan.colour_off[k]; p < plan.colour_off[k + 1]; ++p)
    {
      const int32_t c = plan.cells[p];
      if (c < 0 || c >= ncells || seen[c]) fail("colour plan: cell listed twice or out of range");
      seen[c] = 1;
      for (int t = 0; t < nd; ++t)
      {
        int32_t& s = stamp[tdm[(int64_t)c * nd + t]];
        if (s == k) fail("colour plan: two cells of colour %d share a dof", k);
      }
      for (int t = 0; t < nd; ++t) stamp[tdm[(int64_t)c * nd + t]] = k;
    }
}

void build_brick_plan(int P, int64_t ncells, int64_t ndofs, const int32_t* tdm,
                      const float* centroid, BrickShape brick, int W, int nloc_cap,
                      BrickPlan& plan, const uint8_t* dof_shared, int word_bytes, bool allow_tuned,
                      const int32_t* cell_ijk_in, bool split_parts)
{
  const int n = P + 1, nd = n * n * n;
  if (ndofs > (int64_t)BD_MASK) fail("brick plan: more than 2^30 local dofs");
  if (nloc_cap > 65535) nloc_cap = 65535;
  if (nloc_cap < nd) fail("brick plan: shared-memory dof capacity %d below one cell", nloc_cap);
  if (W < 1) fail("brick plan: W < 1");
  for (int a = 0; a < 3; ++a)
    if (brick.e[a] < 1 || brick.e[a] > 16) fail("brick plan: brick edge %d out of range [1,16]", brick.e[a]);
  const int64_t max_cells = brick.cells();
  plan = BrickPlan();
  plan.P = P;
  plan.nd = nd;
  plan.ndp = (nd + 7) & ~7;
  plan.W = W;
  plan.ncells = ncells;
  plan.ndofs = ndofs;

  const bool verbose = std::getenv("WFX_VERBOSE") && std::atoi(std::getenv("WFX_VERBOSE")) > 1;
  auto tnow = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double tlast = tnow();
  auto lap = [&](const char* what) {
    if (!verbose) return;
    const double t = tnow();
    std::fprintf(stderr, "[wfx plan] %-28s %8.3f s\n", what, t - tlast);
    tlast = t;
  };
  // ---- 1. spatial keys -------------------------------------------------------
  std::vector<uint64_t> key((size_t)ncells);
  std::vector<uint32_t> parity((size_t)ncells, 0);
  std::vector<int32_t> cell_ijk; // integer grid coordinates of the cells (when centroids are given)
  auto set_key = [&](int64_t c, const int64_t (&ia)[3]) {
    uint64_t b[3], loc[3];
    for (int a = 0; a < 3; ++a)
    {
      b[a] = (uint64_t)(ia[a] / brick.e[a]);
      loc[a] = (uint64_t)(ia[a] % brick.e[a]);
      cell_ijk[3 * c + a] = (int32_t)ia[a];
    }
    // brick coordinates in the high bits, in-brick position in the low bits
    key[c] = (b[0] << 44) | (b[1] << 28) | (b[2] << 12) | (loc[0] << 8) | (loc[1] << 4) | loc[2];
    parity[c] = (uint32_t)((b[0] & 1) | ((b[1] & 1) << 1) | ((b[2] & 1) << 2));
  };
  if (cell_ijk_in && ncells > 0)
  {
    cell_ijk.resize((size_t)ncells * 3);
    for (int64_t c = 0; c < ncells; ++c)
    {
      const int64_t ia[3] = {cell_ijk_in[3 * c], cell_ijk_in[3 * c + 1], cell_ijk_in[3 * c + 2]};
      for (int a = 0; a < 3; ++a)
        if (ia[a] < 0 || ia[a] > 0xFFFF) fail("brick plan: cell coordinate out of range");
      set_key(c, ia);
    }
  }
  else if (centroid && ncells > 0)
  {
    cell_ijk.resize((size_t)ncells * 3);
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int64_t c = 0; c < ncells; ++c)
      for (int a = 0; a < 3; ++a)
      {
        lo[a] = std::min(lo[a], (double)centroid[3 * c + a]);
        hi[a] = std::max(hi[a], (double)centroid[3 * c + a]);
      }
    // cell spacing of the (assumed roughly uniform) mesh: centroids of a structured
    // grid span (n_a - 1) h per axis; solve prod (ext_a + h) = ncells h^3 by iteration.
    double ext[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
    double h = std::cbrt(std::max(ext[0], 1e-300) * std::max(ext[1], 1e-300) * std::max(ext[2], 1e-300)
                         / (double)ncells);
    if (!(h > 0)) h = 1.0;
    for (int it = 0; it < 50; ++it)
      h = std::cbrt((ext[0] + h) * (ext[1] + h) * (ext[2] + h) / (double)ncells);
    for (int64_t c = 0; c < ncells; ++c)
    {
      int64_t ia[3];
      for (int a = 0; a < 3; ++a)
      {
        ia[a] = (int64_t)std::floor((centroid[3 * c + a] - lo[a]) / h + 0.5);
        if (ia[a] < 0) ia[a] = 0;
        if (ia[a] > 0xFFFF) ia[a] = 0xFFFF;
      }
      set_key(c, ia);
    }
  }
  else
  {
    for (int64_t c = 0; c < ncells; ++c) key[c] = ((uint64_t)(c / max_cells)) << 12;
  }
  std::vector<int32_t> order((size_t)ncells);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return key[a] < key[b]; });

  lap("keys + sort");
  // ---- 2. batches: runs of equal brick key, capacity-limited --------------------
  struct Batch
  {
    int64_t begin, end; // range in `order`
    int colour = -1;
    uint32_t parity = 0;
    int nloc = 0;
  };
  std::vector<Batch> batches;
  {
    int64_t s = 0;
    while (s < ncells)
    {
      int64_t e = s + 1;
      const uint64_t bk = key[order[s]] >> 12;
      while (e < ncells && (key[order[e]] >> 12) == bk && e - s < max_cells) ++e;
      Batch b;
      b.begin = s;
      b.end = e;
      b.parity = parity[order[s]];
      batches.push_back(b);
      s = e;
    }
  }
  // enforce the shared-memory dof capacity by splitting
  std::vector<int32_t> g2l((size_t)ndofs, -1);
  std::vector<int32_t> uniq;
  auto count_unique = [&](const Batch& b) {
    uniq.clear();
    for (int64_t p = b.begin; p < b.end; ++p)
    {
      const int32_t* d = tdm + (int64_t)order[p] * nd;
      for (int t = 0; t < nd; ++t)
        if (g2l[d[t]] < 0)
        {
          g2l[d[t]] = 1;
          uniq.push_back(d[t]);
        }
    }
    for (int32_t d : uniq) g2l[d] = -1;
    return (int)uniq.size();
  };
  for (size_t i = 0; i < batches.size();)
  {
    const int nl = count_unique(batches[i]);
    if (nl > nloc_cap && batches[i].end - batches[i].begin > 1)
    {
      Batch lo = batches[i], hi = batches[i];
      lo.end = hi.begin = (batches[i].begin + batches[i].end) / 2;
      batches[i] = lo;
      batches.insert(batches.begin() + i + 1, hi);
      continue; // re-examine the lower half
    }
    batches[i].nloc = nl;
    ++i;
  }
  const int nb = (int)batches.size();
  plan.nbatches = nb;

  lap("batches + capacity");
  // ---- 3. batch colouring (greedy; the brick-coordinate parity is tried first, which
  //         gives the optimal 8 colours on structured meshes) ------------------------
  std::vector<uint64_t> dofmask((size_t)ndofs, 0);
  int ncol = 0;
  for (int i = 0; i < nb; ++i)
  {
    uint64_t m = 0;
    for (int64_t p = batches[i].begin; p < batches[i].end; ++p)
    {
      const int32_t* d = tdm + (int64_t)order[p] * nd;
      for (int t = 0; t < nd; ++t) m |= dofmask[d[t]];
    }
    int col = (int)batches[i].parity;
    if ((m >> col) & 1) col = lowest_zero_bit(m);
    if (col < 0) fail("batch colouring needs more than 64 colours");
    const uint64_t bit = 1ull << col;
    for (int64_t p = batches[i].begin; p < batches[i].end; ++p)
    {
      const int32_t* d = tdm + (int64_t)order[p] * nd;
      for (int t = 0; t < nd; ++t) dofmask[d[t]] |= bit;
    }
    batches[i].colour = col;
    ncol = std::max(ncol, col + 1);
  }
  // compact away unused colours, keep order
  {
    std::vector<int> used(ncol, 0), remap(ncol, -1);
    for (auto& b : batches) used[b.colour] = 1;
    int k = 0;
    for (int c = 0; c < ncol; ++c)
      if (used[c]) remap[c] = k++;
    // dofmask bits are remapped lazily below through first/last per dof
    std::vector<uint64_t>().swap(dofmask);
    for (auto& b : batches) b.colour = remap[b.colour];
    ncol = k;
  }
  // Distributed meshes: batches that touch a dof shared with another rank ("interface"
  // batches) run first so that the halo exchange of those dofs can overlap the interior
  // batches.  Execution colour = part * ncol + colour, part 0 = interface, 1 = interior.
  plan.part_split = 0;
  if (dof_shared && split_parts)
  {
    for (auto& b : batches)
    {
      bool iface = false;
      for (int64_t p = b.begin; p < b.end && !iface; ++p)
      {
        const int32_t* d = tdm + (int64_t)order[p] * nd;
        for (int t = 0; t < nd; ++t)
          if (dof_shared[d[t]]) { iface = true; break; }
      }
      if (!iface) b.colour += ncol;
    }
    plan.part_split = ncol;
    ncol *= 2;
  }
  plan.ncolours = ncol;
  std::stable_sort(batches.begin(), batches.end(),
                   [](const Batch& a, const Batch& b) { return a.colour < b.colour; });
  plan.colour_off.assign(ncol + 1, 0);
  for (auto& b : batches) plan.colour_off[b.colour + 1]++;
  for (int c = 0; c < ncol; ++c) plan.colour_off[c + 1] += plan.colour_off[c];

  lap("batch colouring");
  // first / last (execution) colour touching each dof.  Batches of one colour share no dof, so the
  // batches of a colour are processed by several host threads without conflicts.
  std::vector<int16_t> cmin((size_t)ndofs, 32767), cmax((size_t)ndofs, -1);
  for (int col = 0; col < ncol; ++col)
    parallel_for(plan.colour_off[col + 1] - plan.colour_off[col], [&](int64_t q0, int64_t q1) {
      for (int64_t q = q0; q < q1; ++q)
      {
        const Batch& b = batches[plan.colour_off[col] + q];
        for (int64_t p = b.begin; p < b.end; ++p)
        {
          const int32_t* d = tdm + (int64_t)order[p] * nd;
          for (int t = 0; t < nd; ++t)
          {
            cmin[d[t]] = std::min<int16_t>(cmin[d[t]], (int16_t)b.colour);
            cmax[d[t]] = std::max<int16_t>(cmax[d[t]], (int16_t)b.colour);
          }
        }
      }
    }, 64);
  for (int64_t i = 0; i < ndofs; ++i)
    if (cmax[i] < 0) plan.untouched.push_back((int32_t)i);

  lap("first / last colours");
  // ---- 4. per-batch dof lists, rounds and local dofmaps ---------------------------
  // Again colour by colour on several host threads: the batches of a colour touch disjoint entries of
  // the global -> local scratch map.  Every batch fills its own output; the plan arrays are the
  // concatenation in batch order, so the result does not depend on the number of threads.
  plan.dof_off.assign(nb + 1, 0);
  plan.round_off.assign(nb + 1, 0);
  const bool have_coords = !cell_ijk.empty();
  // Off by default: the 7 % padding pushes two resident CTAs over the 196 KB shared-memory
  // carve-out step, L1 shrinks from 60 to 28 KB and the kernel loses 20 % (measured, round 1) --
  // far more than the conflict-free layout gains.  WFX_TUNED_STRIDES=1 enables it for experiments.
  bool want_tuned = false;
  if (const char* ev = std::getenv("WFX_TUNED_STRIDES")) want_tuned = allow_tuned && std::atoi(ev) != 0;
  const Strides lay = have_coords ? find_strides(P, brick, word_bytes, want_tuned) : Strides();
  plan.Sx = lay.Sx;
  plan.Sy = lay.Sy;
  struct BatchOut
  {
    std::vector<uint32_t> bdofs;
    std::vector<int32_t> slot_cell;
    std::vector<uint16_t> slot_base, ldm;
    int nrounds = 0, nslots = 0, n_private = 0;
    uint8_t regular = 0;
  };
  std::vector<BatchOut> outs((size_t)nb);
  const int ndp = plan.ndp;
  auto process = [&](int i, std::vector<int32_t>& uq, std::vector<int>& place, std::vector<uint64_t>& lmask,
                     std::vector<int>& ccol, std::vector<uint8_t>& slot_used) {
    const Batch& b = batches[i];
    BatchOut& o = outs[i];
    const int nc = (int)(b.end - b.begin);
    // unique dofs, ascending
    uq.clear();
    for (int64_t p = b.begin; p < b.end; ++p)
    {
      const int32_t* d = tdm + (int64_t)order[p] * nd;
      for (int t = 0; t < nd; ++t)
        if (g2l[d[t]] < 0)
        {
          g2l[d[t]] = 1;
          uq.push_back(d[t]);
        }
    }
    std::sort(uq.begin(), uq.end());
    int nloc = (int)uq.size();
    if (nloc > 65535) fail("brick plan: batch with %d dofs exceeds 16-bit local index", nloc);
    for (int l = 0; l < nloc; ++l) g2l[uq[l]] = l;
    // Placement of the batch dofs in the shared arrays.  Default: ascending global dof.  If the
    // batch is a regular brick (every dof gets one consistent lattice coordinate from the
    // integer cell coordinates and the tensor index of its points), place dof (X,Y,Z) at
    // X*Sx + Y*Sy + Z with strides searched so that the per-cell gather / scatter of a k-plane
    // (one lane per (i,j)) is free of shared-memory bank conflicts; unused positions are holes.
    place.assign(nloc, -1);
    int nslots = nloc;
    bool regular = false;
    if (have_coords && lay.ok)
    {
      int org[3] = {1 << 30, 1 << 30, 1 << 30};
      for (int64_t p = b.begin; p < b.end; ++p)
        for (int a = 0; a < 3; ++a) org[a] = std::min(org[a], cell_ijk[3 * (int64_t)order[p] + a]);
      regular = true;
      for (int64_t p = b.begin; p < b.end && regular; ++p)
      {
        const int32_t c = order[p];
        const int32_t* d = tdm + (int64_t)c * nd;
        int l3[3];
        for (int a = 0; a < 3; ++a) l3[a] = cell_ijk[3 * (int64_t)c + a] - org[a];
        if (l3[0] >= brick.e[0] || l3[1] >= brick.e[1] || l3[2] >= brick.e[2]) { regular = false; break; }
        for (int k = 0; k < n && regular; ++k)
          for (int ii = 0; ii < n && regular; ++ii)
            for (int jj = 0; jj < n; ++jj)
            {
              const int X = P * l3[0] + ascpos(ii, P), Y = P * l3[1] + ascpos(jj, P), Z = P * l3[2] + ascpos(k, P);
              const int ps = X * lay.Sx + Y * lay.Sy + Z;
              int& cur = place[g2l[d[k * n * n + ii * n + jj]]];
              if (cur >= 0 && cur != ps) { regular = false; break; }
              cur = ps;
            }
      }
      if (regular)
      {
        nslots = P * brick.e[0] * lay.Sx + P * brick.e[1] * lay.Sy + P * brick.e[2] + 1;
        if (nslots > nloc_cap || nslots > 65535) regular = false;
        // positions must be distinct
        if (regular)
        {
          slot_used.assign(nslots, 0);
          for (int l = 0; l < nloc && regular; ++l)
          {
            if (place[l] < 0 || place[l] >= nslots || slot_used[place[l]]) regular = false;
            else slot_used[place[l]] = 1;
          }
        }
      }
      if (!regular)
      {
        nslots = nloc;
        place.assign(nloc, -1);
      }
    }
    o.regular = regular ? 1 : 0;
    if (place.empty() || (nloc > 0 && place[0] < 0))
      for (int l = 0; l < nloc; ++l) place[l] = l;
    o.nslots = nslots;
    o.bdofs.assign((size_t)nslots, BD_HOLE);
    for (int l = 0; l < nloc; ++l)
    {
      const int32_t d = uq[l];
      uint32_t e = (uint32_t)d;
      if (cmin[d] == b.colour) e |= BD_FIRST;
      // shared dofs are complete only after the halo sum: never LAST (no fused scaling) here
      if (cmax[d] == b.colour && !(dof_shared && dof_shared[d])) e |= BD_LAST;
      if ((e & BD_FIRST) && (e & BD_LAST)) o.n_private++;
      o.bdofs[place[l]] = e;
      g2l[d] = place[l]; // local index used by the cells = position in the shared arrays
    }
    nloc = nslots;
    // colour the batch's cells so that cells of one round share no local dof
    lmask.assign(nloc, 0);
    ccol.assign(nc, 0);
    int nlc = 0;
    for (int q = 0; q < nc; ++q)
    {
      const int32_t* d = tdm + (int64_t)order[b.begin + q] * nd;
      uint64_t m = 0;
      for (int t = 0; t < nd; ++t) m |= lmask[g2l[d[t]]];
      const int col = lowest_zero_bit(m);
      if (col < 0) fail("in-batch cell colouring needs more than 64 colours");
      for (int t = 0; t < nd; ++t) lmask[g2l[d[t]]] |= 1ull << col;
      ccol[q] = col;
      nlc = std::max(nlc, col + 1);
    }
    for (int col = 0; col < nlc; ++col)
    {
      int filled = 0;
      for (int q = 0; q < nc; ++q)
      {
        if (ccol[q] != col) continue;
        if (filled == 0)
        {
          o.slot_cell.resize(o.slot_cell.size() + W, -1);
          o.slot_base.resize(o.slot_base.size() + W, 0);
          o.ldm.resize(o.ldm.size() + (size_t)W * ndp, 0);
          ++o.nrounds;
        }
        const size_t slot = o.slot_cell.size() - W + filled;
        const int32_t cell = order[b.begin + q];
        o.slot_cell[slot] = cell;
        const int32_t* d = tdm + (int64_t)cell * nd;
        for (int t = 0; t < nd; ++t) o.ldm[slot * ndp + t] = (uint16_t)g2l[d[t]];
        // position of the cell's origin corner (point i=j=k=0): for a regular brick every other
        // point of the cell is at a fixed offset from it (ascpos(i)*Sx + ascpos(j)*Sy + ascpos(k))
        o.slot_base[slot] = (uint16_t)g2l[d[0]];
        filled = (filled + 1) % W;
      }
    }
    for (int32_t d : uq) g2l[d] = -1;
  };
  for (int col = 0; col < ncol; ++col)
    parallel_for(plan.colour_off[col + 1] - plan.colour_off[col], [&](int64_t q0, int64_t q1) {
      std::vector<int32_t> uq;
      std::vector<int> place, ccol;
      std::vector<uint64_t> lmask;
      std::vector<uint8_t> slot_used;
      for (int64_t q = q0; q < q1; ++q) process(plan.colour_off[col] + (int)q, uq, place, lmask, ccol, slot_used);
    }, 64);
  lap("per-batch lists (threads)");
  // concatenate in batch order
  {
    size_t nbd = 0, nsl = 0;
    for (const BatchOut& o : outs) nbd += o.bdofs.size(), nsl += o.slot_cell.size();
    plan.bdofs.reserve(nbd + 4);
    plan.slot_cell.reserve(nsl);
    plan.slot_base.reserve(nsl);
    plan.ldm.reserve(nsl * (size_t)ndp);
    plan.batch_regular.reserve((size_t)nb);
    for (int i = 0; i < nb; ++i)
    {
      BatchOut& o = outs[i];
      plan.bdofs.insert(plan.bdofs.end(), o.bdofs.begin(), o.bdofs.end());
      plan.slot_cell.insert(plan.slot_cell.end(), o.slot_cell.begin(), o.slot_cell.end());
      plan.slot_base.insert(plan.slot_base.end(), o.slot_base.begin(), o.slot_base.end());
      plan.ldm.insert(plan.ldm.end(), o.ldm.begin(), o.ldm.end());
      plan.dof_off[i + 1] = (int64_t)plan.bdofs.size();
      plan.round_off[i + 1] = plan.round_off[i] + o.nrounds;
      plan.rounds_max = std::max(plan.rounds_max, o.nrounds);
      plan.nloc_max = std::max(plan.nloc_max, o.nslots);
      plan.n_private += o.n_private;
      plan.n_regular += o.regular;
      plan.batch_regular.push_back(have_coords && lay.ok ? o.regular : 0);
      BatchOut().bdofs.swap(o.bdofs); // release as we go
      std::vector<uint16_t>().swap(o.ldm);
    }
  }
  lap("concatenate");
  // write-back dependencies (see wfx_plan.h): one pass over the dof lists in execution order
  {
    std::vector<int32_t> last_toucher((size_t)ndofs, -1);
    plan.dep_off.assign(nb + 1, 0);
    plan.dep_ids.clear();
    std::vector<int32_t> deps;
    for (int i = 0; i < nb; ++i)
    {
      deps.clear();
      for (int64_t l = plan.dof_off[i]; l < plan.dof_off[i + 1]; ++l)
      {
        const uint32_t e = plan.bdofs[l];
        if (e == BD_HOLE) continue;
        int32_t& lt = last_toucher[e & BD_MASK];
        if (lt >= 0 && (deps.empty() || deps.back() != lt)) deps.push_back(lt);
        lt = i;
      }
      std::sort(deps.begin(), deps.end());
      deps.erase(std::unique(deps.begin(), deps.end()), deps.end());
      plan.dep_ids.insert(plan.dep_ids.end(), deps.begin(), deps.end());
      plan.dep_off[i + 1] = (int32_t)plan.dep_ids.size();
    }
  }
  lap("write-back dependencies");
  plan.nrounds_total = plan.round_off[nb];
  plan.n_slots_padded = plan.nrounds_total * W - ncells;
  // the tuned strides only pay off when every batch runs the regular-brick kernel; otherwise
  // their padding just costs shared memory: plan again with the compact strides
  if (lay.tuned && plan.n_regular != nb)
  {
    BrickPlan compact;
    build_brick_plan(P, ncells, ndofs, tdm, centroid, brick, W, nloc_cap, compact, dof_shared, word_bytes, false,
                     cell_ijk_in, split_parts);
    plan = std::move(compact);
  }
}

void verify_brick_plan(const BrickPlan& plan, const int32_t* tdm, const uint8_t* dof_shared)
{
  const int nd = plan.nd, W = plan.W, nb = plan.nbatches;
  const int64_t ncells = plan.ncells, ndofs = plan.ndofs;
  if ((int)plan.colour_off.size() != plan.ncolours + 1 || plan.colour_off.back() != nb)
    fail("brick plan: colour offsets inconsistent");
  std::vector<uint8_t> cell_seen((size_t)ncells, 0);
  std::vector<int32_t> colour_stamp((size_t)ndofs, -1); // last colour that touched the dof
  std::vector<uint8_t> nfirst((size_t)ndofs, 0), nlast((size_t)ndofs, 0), touched((size_t)ndofs, 0);
  std::vector<int32_t> prev_colour((size_t)ndofs, -1);
  std::vector<int32_t> round_stamp;
  int64_t cells_total = 0;
  for (int col = 0; col < plan.ncolours; ++col)
    for (int b = plan.colour_off[col]; b < plan.colour_off[col + 1]; ++b)
    {
      const int64_t d0 = plan.dof_off[b], d1 = plan.dof_off[b + 1];
      const int nloc = (int)(d1 - d0);
      if (nloc > plan.nloc_max) fail("brick plan: nloc_max too small");
      for (int64_t l = d0; l < d1; ++l)
      {
        const uint32_t e = plan.bdofs[l];
        if (e == BD_HOLE) continue; // unused position of a regular-brick placement
        const int32_t d = (int32_t)(e & BD_MASK);
        if (d < 0 || d >= ndofs) fail("brick plan: dof out of range");
        if (colour_stamp[d] == col) fail("brick plan: two batches of colour %d share dof %d", col, d);
        if (dof_shared && dof_shared[d] && col >= plan.part_split && plan.part_split > 0)
          fail("brick plan: interior batch touches shared dof %d", d);
        colour_stamp[d] = col;
        const bool first = e & BD_FIRST, last = e & BD_LAST;
        if (first && touched[d]) fail("brick plan: FIRST flag on an already touched dof");
        if (!first && !touched[d]) fail("brick plan: dof %d first touched without FIRST flag", d);
        if (nlast[d]) fail("brick plan: dof %d touched after its LAST batch", d);
        touched[d] = 1;
        nfirst[d] += first;
        nlast[d] += last;
      }
      round_stamp.assign(nloc, -1);
      for (int r = plan.round_off[b]; r < plan.round_off[b + 1]; ++r)
        for (int s = 0; s < W; ++s)
        {
          const int64_t slot = (int64_t)r * W + s;
          const int32_t cell = plan.slot_cell[slot];
          if (cell < 0) continue;
          if (cell >= ncells || cell_seen[cell]) fail("brick plan: cell out of range or listed twice");
          cell_seen[cell] = 1;
          ++cells_total;
          for (int t = 0; t < nd; ++t)
          {
            const int l = plan.ldm[slot * plan.ndp + t];
            if (l >= nloc) fail("brick plan: local dof index out of range");
            if ((int32_t)(plan.bdofs[d0 + l] & BD_MASK) != tdm[(int64_t)cell * nd + t])
              fail("brick plan: local dofmap does not reproduce the dofmap");
            if (round_stamp[l] == r) fail("brick plan: two cells of one round share a dof");
          }
          for (int t = 0; t < nd; ++t) round_stamp[plan.ldm[slot * plan.ndp + t]] = r;
          // regular bricks: the kernels compute positions arithmetically from the cell's base
          if (!plan.batch_regular.empty() && plan.batch_regular[b])
          {
            const int n1 = plan.P + 1;
            for (int k = 0; k < n1; ++k)
              for (int i = 0; i < n1; ++i)
                for (int j = 0; j < n1; ++j)
                  if (plan.ldm[slot * plan.ndp + (k * n1 + i) * n1 + j]
                      != plan.slot_base[slot] + ascpos(i, plan.P) * plan.Sx + ascpos(j, plan.P) * plan.Sy + ascpos(k, plan.P))
                    fail("brick plan: regular batch with a non-arithmetic position");
          }
        }
    }
  if (cells_total != ncells) fail("brick plan: %lld of %lld cells covered", (long long)cells_total, (long long)ncells);
  // write-back dependencies: only on earlier batches, and every pair of batches sharing a dof is
  // ordered through the chain of last touchers
  if (!plan.dep_off.empty())
  {
    if ((int)plan.dep_off.size() != nb + 1) fail("brick plan: dependency offsets inconsistent");
    std::vector<int32_t> lt((size_t)ndofs, -1);
    for (int b = 0; b < nb; ++b)
    {
      for (int32_t p = plan.dep_off[b]; p < plan.dep_off[b + 1]; ++p)
        if (plan.dep_ids[p] < 0 || plan.dep_ids[p] >= b) fail("brick plan: batch %d depends on a later batch", b);
      for (int64_t l = plan.dof_off[b]; l < plan.dof_off[b + 1]; ++l)
      {
        const uint32_t e = plan.bdofs[l];
        if (e == BD_HOLE) continue;
        const int32_t prev = lt[e & BD_MASK];
        if (prev >= 0
            && !std::binary_search(plan.dep_ids.begin() + plan.dep_off[b], plan.dep_ids.begin() + plan.dep_off[b + 1], prev))
          fail("brick plan: batch %d misses its dependency on batch %d", b, prev);
        lt[e & BD_MASK] = b;
      }
    }
  }
  int64_t nunt = 0;
  for (int64_t d = 0; d < ndofs; ++d)
  {
    if (!touched[d]) { ++nunt; continue; }
    const int want_last = (dof_shared && dof_shared[d]) ? 0 : 1;
    if (nfirst[d] != 1 || nlast[d] != want_last) fail("brick plan: dof %lld has %d FIRST / %d LAST marks", (long long)d, nfirst[d], nlast[d]);
  }
  if (nunt != (int64_t)plan.untouched.size()) fail("brick plan: untouched list inconsistent");
}
} // namespace wfx

// Debug entry point (not part of the drop-in boundary): builds both plans for a dofmap on
// the host, verifies every invariant and reports statistics.  Needs no GPU.
//   stats[0]=cell colours  [1]=batches  [2]=batch colours  [3]=nloc_max  [4]=rounds
//   [5]=padded slots  [6]=private (first&last) batch dofs  [7]=total batch dofs
//   [8]=untouched dofs
// Debug entry point: structured_cell_coords on host arrays.  *ok = 1 and ijk_out [ncells][3] filled
// when the mesh is one consistently oriented structured block, else *ok = 0.
extern "C" int wfx_debug_structured_coords(int64_t ncells, int64_t npts, const int32_t* xdofs_host,
                                           int32_t* ijk_out, int* ok)
{
  WFX_API_BEGIN
  using namespace wfx;
  if (!xdofs_host || !ok) fail("NULL argument");
  for (int64_t q = 0; q < ncells * 8; ++q)
    if (xdofs_host[q] < 0 || xdofs_host[q] >= npts) fail("geometry dofmap entry out of range");
  std::vector<int32_t> ijk;
  *ok = structured_cell_coords(ncells, npts, xdofs_host, ijk) ? 1 : 0;
  if (*ok && ijk_out) std::memcpy(ijk_out, ijk.data(), ijk.size() * sizeof(int32_t));
  WFX_API_END
}

namespace wfx
{
void detect_axis_perm(int n, int64_t ncells, const int32_t* tdm, int (&s)[3])
{
  const int n2 = n * n, nd = n2 * n;
  int64_t votes[3] = {0, 0, 0};
  const int64_t ns = std::min<int64_t>(ncells, 256), step = std::max<int64_t>(1, ncells / std::max<int64_t>(ns, 1));
  int64_t seen = 0;
  for (int64_t c = 0; c < ncells; c += step, ++seen)
  {
    // tensor point (i,j,k) is entry k*n2 + i*n + j; 1-D index 0 is lattice position 0, index 2 position 1
    const int32_t* d = tdm + c * nd;
    const int32_t o = d[0];
    if (std::abs(d[2 * n + 0] - o) == 1) ++votes[0]; // (2,0,0)
    if (std::abs(d[2] - o) == 1) ++votes[1];         // (0,2,0)
    if (std::abs(d[2 * n2] - o) == 1) ++votes[2];    // (0,0,2)
  }
  s[0] = 0, s[1] = 1, s[2] = 2;
  if (2 * votes[2] > seen) s[0] = 1, s[1] = 2, s[2] = 0;
  else if (2 * votes[0] > seen) s[0] = 2, s[1] = 0, s[2] = 1;
}

void build_stream_plan(int P, int64_t ncells, int64_t ndofs, const int32_t* tdm, const float* centroid,
                       const int32_t* cell_ijk, bool brick_order, BrickShape brick, int W,
                       const uint8_t* dof_shared, bool split_parts, bool relabel_axes, StreamPlan& sp)
{
  const int n = P + 1, n2 = n * n, nd = n2 * n;
  if (n < 3) fail("stream plan: degree %d not supported", P);
  if (ndofs > (int64_t)BD_MASK) fail("stream plan: more than 2^30 local dofs");
  if (dof_shared && !brick_order) fail("stream plan: colour-ordered cells do not take partitioned meshes");
  sp = StreamPlan();
  sp.brick_order = brick_order;
  sp.P = P;
  sp.nd = nd;
  sp.ncells = ncells;
  sp.ndofs = ndofs;
  // key[c]: position of cell c in the execution order (launch, iteration)
  std::vector<uint32_t> key((size_t)ncells, 0);
  if (brick_order)
  {
    BrickPlan bp;
    build_brick_plan(P, ncells, ndofs, tdm, centroid, brick, W, 65535, bp, dof_shared, 8, false, cell_ijk, split_parts);
    sp.W = bp.W;
    sp.ncolours = bp.ncolours;
    sp.part_split = bp.part_split;
    sp.nbatches = bp.nbatches;
    sp.rounds_max = bp.rounds_max;
    sp.colour_off = bp.colour_off;
    bool ur = bp.nbatches > 0;
    for (int b = 0; b < bp.nbatches; ++b) ur = ur && bp.round_off[b + 1] - bp.round_off[b] == bp.round_off[1];
    sp.uni_nr = ur ? bp.round_off[1] : 0;
    if (bp.ncolours >= 65536 || bp.rounds_max >= 65536) fail("stream plan: order key overflow");
    for (int k = 0; k < bp.ncolours; ++k)
      for (int b = bp.colour_off[k]; b < bp.colour_off[k + 1]; ++b)
        for (int r = bp.round_off[b]; r < bp.round_off[b + 1]; ++r)
          for (int w = 0; w < bp.W; ++w)
          {
            const int32_t c = bp.slot_cell[(size_t)r * bp.W + w];
            if (c >= 0) key[c] = (uint32_t)k * 65536u + (uint32_t)(r - bp.round_off[b]);
          }
    sp.round_off = std::move(bp.round_off);
    sp.slot_cell = std::move(bp.slot_cell);
  }
  else
  {
    CellColourPlan cp;
    build_cell_colour_plan(nd, ncells, ndofs, tdm, cp);
    sp.ncolours = cp.ncolours;
    sp.colour_off = cp.colour_off;
    for (int k = 0; k < cp.ncolours; ++k)
      for (int32_t p = cp.colour_off[k]; p < cp.colour_off[k + 1]; ++p) key[cp.cells[p]] = (uint32_t)k;
    sp.cells = std::move(cp.cells);
  }
  std::vector<uint32_t> kmin((size_t)ndofs, 0xffffffffu), kmax((size_t)ndofs, 0);
  for (int64_t c = 0; c < ncells; ++c)
    for (int t = 0; t < nd; ++t)
    {
      const int32_t d = tdm[c * nd + t];
      kmin[d] = std::min(kmin[d], key[c]);
      kmax[d] = std::max(kmax[d], key[c]);
    }
  if (relabel_axes && ncells > 0) detect_axis_perm(n, ncells, tdm, sp.axis_perm);
  const int s0 = sp.axis_perm[0], s1 = sp.axis_perm[1], s2 = sp.axis_perm[2];
  sp.tdmf.resize((size_t)ncells * nd);
  parallel_for(ncells, [&](int64_t cb, int64_t ce) {
    for (int64_t c = cb; c < ce; ++c)
      for (int kp = 0; kp < n; ++kp)
        for (int ip = 0; ip < n; ++ip)
          for (int jp = 0; jp < n; ++jp)
          {
            int sidx[3];
            sidx[s0] = ip, sidx[s1] = jp, sidx[s2] = kp; // the point's indices on the mesh's own axes
            const int32_t d = tdm[c * nd + sidx[2] * n2 + sidx[0] * n + sidx[1]];
            uint32_t e = (uint32_t)d;
            if (key[c] == kmin[d]) e |= BD_FIRST;
            if (key[c] == kmax[d] && !(dof_shared && dof_shared[d])) e |= BD_LAST;
            sp.tdmf[c * nd + kp * n2 + ip * n + jp] = e;
          }
  });
  for (int64_t d = 0; d < ndofs; ++d)
    if (kmin[d] == 0xffffffffu) sp.untouched.push_back((int32_t)d);
}

void verify_stream_plan(const StreamPlan& sp, const int32_t* tdm, const uint8_t* dof_shared)
{
  const int nd = sp.nd, n = sp.P + 1, n2 = n * n;
  if ((int64_t)sp.tdmf.size() != sp.ncells * nd) fail("stream plan: wrong dofmap size");
  {
    int seen[3] = {0, 0, 0};
    for (int a = 0; a < 3; ++a)
    {
      if (sp.axis_perm[a] < 0 || sp.axis_perm[a] > 2) fail("stream plan: bad axis relabelling");
      seen[sp.axis_perm[a]]++;
    }
    if (seen[0] != 1 || seen[1] != 1 || seen[2] != 1) fail("stream plan: axis relabelling is not a permutation");
  }
  // replay the execution order: step = (launch, iteration); stamp[d] = last step that touched dof d
  std::vector<int64_t> stamp((size_t)sp.ndofs, -1);
  std::vector<uint8_t> done((size_t)sp.ncells, 0), has_last((size_t)sp.ndofs, 0);
  std::vector<int32_t> last_cell((size_t)sp.ndofs, -1), last_pt((size_t)sp.ndofs, -1);
  int64_t step = 0, ndone = 0;
  auto visit = [&](int32_t c) {
    if (c < 0 || c >= sp.ncells || done[c]) fail("stream plan: cell listed twice or out of range");
    done[c] = 1;
    ++ndone;
    const int s0 = sp.axis_perm[0], s1 = sp.axis_perm[1], s2 = sp.axis_perm[2];
    for (int kp = 0; kp < n; ++kp)
      for (int ip = 0; ip < n; ++ip)
        for (int jp = 0; jp < n; ++jp)
        {
          int sidx[3];
          sidx[s0] = ip, sidx[s1] = jp, sidx[s2] = kp;
          const int t = kp * n2 + ip * n + jp;
          const uint32_t e = sp.tdmf[(int64_t)c * nd + t];
          const int32_t d = (int32_t)(e & BD_MASK);
          if (d != tdm[(int64_t)c * nd + sidx[2] * n2 + sidx[0] * n + sidx[1]]) fail("stream plan: dofmap entry does not match the mesh");
          if (stamp[d] == step) fail("stream plan: two cells of one step share dof %d", d);
          if (((e & BD_FIRST) != 0) != (stamp[d] < 0)) fail("stream plan: FIRST flag wrong at dof %d", d);
          if (has_last[d]) fail("stream plan: dof %d touched after its LAST point", d);
          if (e & BD_LAST)
          {
            if (dof_shared && dof_shared[d]) fail("stream plan: LAST set on a rank-shared dof");
            has_last[d] = 1;
          }
          stamp[d] = step;
        }
  };
  if (sp.brick_order)
  {
    if ((int)sp.colour_off.size() != sp.ncolours + 1) fail("stream plan: bad colour offsets");
    for (int k = 0; k < sp.ncolours; ++k)
    {
      for (int b = sp.colour_off[k]; b < sp.colour_off[k + 1]; ++b)
        for (int r = sp.round_off[b]; r < sp.round_off[b + 1]; ++r)
        {
          for (int w = 0; w < sp.W; ++w)
          {
            const int32_t c = sp.slot_cell[(size_t)r * sp.W + w];
            if (c >= 0) visit(c);
          }
          ++step;
        }
    }
    // batches of one launch run concurrently and their rounds interleave arbitrarily: a dof must not be
    // shared between two batches of a launch at all
    std::vector<int32_t> batch_of((size_t)sp.ndofs, -1), colour_of((size_t)sp.ndofs, -1);
    for (int k = 0; k < sp.ncolours; ++k)
      for (int b = sp.colour_off[k]; b < sp.colour_off[k + 1]; ++b)
        for (int r = sp.round_off[b]; r < sp.round_off[b + 1]; ++r)
          for (int w = 0; w < sp.W; ++w)
          {
            const int32_t c = sp.slot_cell[(size_t)r * sp.W + w];
            if (c < 0) continue;
            for (int t = 0; t < nd; ++t)
            {
              const int32_t d = tdm[(int64_t)c * nd + t];
              if (colour_of[d] == k && batch_of[d] != b) fail("stream plan: two batches of launch %d share dof %d", k, d);
              colour_of[d] = k, batch_of[d] = b;
            }
          }
  }
  else
  {
    for (int k = 0; k < sp.ncolours; ++k)
    {
      for (int32_t p = sp.colour_off[k]; p < sp.colour_off[k + 1]; ++p) visit(sp.cells[p]);
      ++step;
    }
  }
  if (ndone != sp.ncells) fail("stream plan: %lld of %lld cells scheduled", (long long)ndone, (long long)sp.ncells);
  for (int64_t d = 0; d < sp.ndofs; ++d)
    if (stamp[d] >= 0 && !has_last[d] && !(dof_shared && dof_shared[d])) fail("stream plan: dof %lld has no LAST point", (long long)d);
  for (int32_t d : sp.untouched)
    if (stamp[d] >= 0) fail("stream plan: dof %d listed as untouched", d);
}
} // namespace wfx

// CPU-callable check of the streamed-cell plan (tests): builds it for a DOLFINx-ordered dofmap and runs
// verify_stream_plan.  out: axis_perm[3]; stats: ncolours, nbatches, part_split, uni_nr, untouched.
extern "C" int wfx_debug_stream_plan_check(int P, int64_t ncells, int64_t ndofs, const int32_t* dofmap_host,
                                           const float* centroid_host, int brick_order, int bx, int by, int bz,
                                           int W, const uint8_t* dof_shared, int relabel_axes, int* axis_perm,
                                           int64_t* stats)
{
  WFX_API_BEGIN
  using namespace wfx;
  const int n = P + 1, n2 = n * n, nd = n2 * n;
  std::vector<int32_t> perm(nd);
  tensor_perm(P, perm.data());
  std::vector<int32_t> tdm((size_t)ncells * nd);
  for (int64_t c = 0; c < ncells; ++c)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j)
        for (int k = 0; k < n; ++k)
        {
          const int32_t d = dofmap_host[c * nd + perm[(i * n + j) * n + k]];
          if (d < 0 || d >= ndofs) fail("dofmap entry out of range");
          tdm[c * nd + k * n2 + i * n + j] = d;
        }
  StreamPlan sp;
  build_stream_plan(P, ncells, ndofs, tdm.data(), centroid_host, nullptr, brick_order != 0, BrickShape(bx, by, bz), W,
                    dof_shared, true, relabel_axes != 0, sp);
  verify_stream_plan(sp, tdm.data(), dof_shared);
  if (axis_perm)
    for (int a = 0; a < 3; ++a) axis_perm[a] = sp.axis_perm[a];
  if (stats)
  {
    stats[0] = sp.ncolours;
    stats[1] = sp.nbatches;
    stats[2] = sp.part_split;
    stats[3] = sp.uni_nr;
    stats[4] = (int64_t)sp.untouched.size();
  }
  WFX_API_END
}

extern "C" int wfx_debug_plan_stats(int P, int64_t ncells, int64_t ndofs,
                                    const int32_t* dofmap_host, const float* centroid_host,
                                    int brick_edge, int W, int nloc_cap, int64_t* stats)
{
  WFX_API_BEGIN
  using namespace wfx;
  const int n = P + 1, n2 = n * n, nd = n2 * n;
  std::vector<int32_t> perm(nd);
  tensor_perm(P, perm.data());
  std::vector<int32_t> tdm((size_t)ncells * nd);
  for (int64_t c = 0; c < ncells; ++c)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j)
        for (int k = 0; k < n; ++k)
        {
          const int32_t d = dofmap_host[c * nd + perm[(i * n + j) * n + k]];
          if (d < 0 || d >= ndofs) fail("dofmap entry out of range");
          tdm[c * nd + k * n2 + i * n + j] = d;
        }
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t0 = now();
  CellColourPlan cp;
  build_cell_colour_plan(nd, ncells, ndofs, tdm.data(), cp);
  const double t_cc = now() - t0;
  verify_cell_colour_plan(cp, nd, ncells, ndofs, tdm.data());
  BrickPlan bp;
  // brick_edge > 255: a non-cubic brick packed as ex | ey << 8 | ez << 16
  const BrickShape shape = brick_edge > 255
                               ? BrickShape(brick_edge & 255, (brick_edge >> 8) & 255, (brick_edge >> 16) & 255)
                               : BrickShape(brick_edge);
  t0 = now();
  build_brick_plan(P, ncells, ndofs, tdm.data(), centroid_host, shape, W, nloc_cap, bp);
  const double t_bp = now() - t0;
  verify_brick_plan(bp, tdm.data());
  if (stats)
  {
    stats[0] = cp.ncolours;
    stats[1] = bp.nbatches;
    stats[2] = bp.ncolours;
    stats[3] = bp.nloc_max;
    stats[4] = bp.nrounds_total;
    stats[5] = bp.n_slots_padded;
    stats[6] = bp.n_private;
    stats[7] = (int64_t)bp.bdofs.size();
    stats[8] = (int64_t)bp.untouched.size();
    stats[9] = bp.n_regular;
    stats[10] = bp.Sx;
    stats[11] = bp.Sy;
    stats[12] = (int64_t)(t_cc * 1e3); // ms: cell colour plan
    stats[13] = (int64_t)(t_bp * 1e3); // ms: brick plan
  }
  WFX_API_END
}
