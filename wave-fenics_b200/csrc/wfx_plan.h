// Host-side execution plans for the atomic-free stiffness scatter.
//
// The reference scatters with atomicAdd (common/cuda/scatter.cu:38-45).  Here the
// scatter is made race-free by construction at setup time:
//
//  CellColourPlan  cells greedily coloured so that no two cells of one colour share a
//                  dof; one launch per colour, plain read-modify-write on y.
//  BrickPlan       cells grouped into spatially compact batches ("bricks").  One CTA
//                  owns one batch: it stages the batch's unique dofs in shared memory,
//                  processes the batch's cells in conflict-free rounds accumulating in
//                  shared memory, and writes every batch dof back exactly once.  Batches
//                  are coloured so that batches of one colour share no dof; colours run
//                  as consecutive launches.  For every dof the lowest-colour batch that
//                  touches it is marked FIRST (it may overwrite instead of accumulate)
//                  and the highest LAST (it may apply the fused diagonal scaling).
#pragma once

#include <cstdint>
#include <functional>
#include <vector>

namespace wfx
{
constexpr uint32_t BD_FIRST = 1u << 30;
constexpr uint32_t BD_LAST = 1u << 31;
constexpr uint32_t BD_MASK = (1u << 30) - 1;
constexpr uint32_t BD_HOLE = BD_MASK; // unused position in a batch's dof list (no flags set)

struct CellColourPlan
{
  int ncolours = 0;
  std::vector<int32_t> colour_off; // [ncolours+1] into cells
  std::vector<int32_t> cells;      // cell ids sorted by colour
};

struct BrickPlan
{
  int P = 0, nd = 0, W = 0; // W = cell slots processed concurrently per round
  int ndp = 0;              // slot stride of ldm: nd rounded up to 8 entries (16 bytes)
  int rounds_max = 0;       // most rounds in one batch
  int64_t ncells = 0, ndofs = 0;
  int nbatches = 0, ncolours = 0;  // ncolours = execution colours (2x the graph colours when split)
  int part_split = 0;              // first execution colour of the interior part (0 = no split)
  int nloc_max = 0;                // largest number of unique dofs in a batch
  int64_t nrounds_total = 0;
  // batches are stored sorted by colour
  std::vector<int32_t> colour_off; // [ncolours+1] into batches
  std::vector<int64_t> dof_off;    // [nbatches+1] into bdofs
  std::vector<uint32_t> bdofs;     // global dof | BD_FIRST | BD_LAST, ascending per batch
  std::vector<int32_t> round_off;  // [nbatches+1] into rounds
  std::vector<int32_t> slot_cell;  // [nrounds_total*W] cell id or -1
  std::vector<uint16_t> slot_base; // [nrounds_total*W] position of the cell's origin corner in the batch arrays
  std::vector<uint16_t> ldm;       // [nrounds_total*W][ndp] batch-local dof, k-major point order
  std::vector<int32_t> untouched;  // vector entries no cell references
  // Write-back dependencies for the single-launch (persistent) kernel: batch b may add to y only
  // after the batches dep_ids[dep_off[b] .. dep_off[b+1]) have written back -- for every dof of b the
  // batch that touched it last before b (earlier in the colour-sorted order; its own predecessors are
  // covered transitively)
  std::vector<int32_t> dep_off, dep_ids;
  // statistics
  int64_t n_slots_padded = 0;
  int64_t n_private = 0;           // bdofs entries that are FIRST and LAST
  int n_regular = 0;               // batches placed as regular bricks (bank-conflict-free layout)
  std::vector<uint8_t> batch_regular; // [nbatches] 1 if the batch is a regular brick
  int Sx = 0, Sy = 0;              // strides of that placement
};

// Cells per brick along the three grid axes (x slowest).  A plain int means a cube.
struct BrickShape
{
  int e[3];
  BrickShape(int cube) : e{cube, cube, cube} {}
  BrickShape(int ex, int ey, int ez) : e{ex, ey, ez} {}
  int64_t cells() const { return (int64_t)e[0] * e[1] * e[2]; }
  bool cubic() const { return e[0] == e[1] && e[1] == e[2]; }
};

// Integer grid coordinates of the cells of a structured hexahedral mesh from its CONNECTIVITY
// (geometry dofmap, vertex v = ix + 2 iy + 4 iz): face neighbours are found through shared
// vertices, orientations must agree, coordinates follow by breadth-first search.  Independent of the
// geometry, so sheared, graded or curved structured meshes get exact coordinates.  Returns false (ijk
// untouched) if the mesh is not one consistently oriented structured block.
bool structured_cell_coords(int64_t ncells, int64_t npts, const int32_t* xdofs, std::vector<int32_t>& ijk);

// runs fn(begin, end) on slices of [0, n) with the host's hardware threads
void parallel_for(int64_t n, const std::function<void(int64_t, int64_t)>& fn, int64_t min_parallel = 4096);

// tdm: tensor-ordered dofmap in the kernels' k-major point order, [ncells][nd]
void build_cell_colour_plan(int nd, int64_t ncells, int64_t ndofs, const int32_t* tdm,
                            CellColourPlan& plan);

// centroid: [ncells][3] or nullptr (then cells are batched in the given order).
// dof_shared: [ndofs] flags of dofs that also live on another rank: never marked LAST (their scaling
// waits for the ghost reduction); with split_parts the batches touching them get their own, earlier
// execution colours (interface part) so that the reduction can overlap the interior part.
// cell_ijk: [ncells][3] exact integer grid coordinates (structured_cell_coords) or nullptr; preferred
// over coordinates estimated from the centroids (which assume a roughly uniform axis-aligned grid).
// brick: cells per brick along each axis (batch capacity = their product); nloc_cap: capacity of
// the shared-memory dof arrays.
void build_brick_plan(int P, int64_t ncells, int64_t ndofs, const int32_t* tdm,
                      const float* centroid, BrickShape brick, int W, int nloc_cap,
                      BrickPlan& plan, const uint8_t* dof_shared = nullptr, int word_bytes = 8,
                      bool allow_tuned = true, const int32_t* cell_ijk = nullptr, bool split_parts = true);

// Execution plan of the streamed-cell kernel (stiff_cell2_kernel): the cells in launch / iteration order
// and the per-point dofmap with FIRST / LAST flags.
//  brick order : launches = execution colours of a brick plan, one CTA per batch, iteration = round;
//  colour order: launches = cell colours, a slot walks consecutive cells of its colour.
// Cells of one (launch, iteration) share no dof -- in brick order within a batch, batches of one launch
// being disjoint anyway.  FIRST marks the point whose cell comes first in that order at its dof, LAST the
// last one (never set on dofs that also live on another rank).
struct StreamPlan
{
  bool brick_order = true;
  int P = 0, nd = 0, W = 0;
  int64_t ncells = 0, ndofs = 0;
  int ncolours = 0, part_split = 0, nbatches = 0, rounds_max = 0;
  int uni_nr = 0;                  // > 0: every batch has this many rounds
  std::vector<int32_t> colour_off; // [ncolours+1] into batches (brick order) or cells (colour order)
  std::vector<int32_t> round_off;  // brick order: [nbatches+1]
  std::vector<int32_t> slot_cell;  // brick order: [nrounds_total*W] cell id or -1
  std::vector<int32_t> cells;      // colour order: cell ids sorted by colour
  int axis_perm[3] = {0, 1, 2};    // kernel axis a' is the mesh's tensor axis axis_perm[a']
  std::vector<uint32_t> tdmf;      // [ncells][nd], point order k'*n^2 + i'*n + j' of the kernel axes: dof | flags
  std::vector<int32_t> untouched;  // vector entries no cell references
};

// Which tensor axis of the cells runs along consecutive dof numbers (lexicographic numberings of
// structured meshes: the fastest grid axis)?  The streamed-cell kernel wants that axis on its fast lane
// index (j'), so that a gather / update instruction touches few sectors: s[1] = that axis, or the
// identity when no axis is contiguous (unstructured or renumbered dofs).
void detect_axis_perm(int n, int64_t ncells, const int32_t* tdm, int (&s)[3]);

void build_stream_plan(int P, int64_t ncells, int64_t ndofs, const int32_t* tdm, const float* centroid,
                       const int32_t* cell_ijk, bool brick_order, BrickShape brick, int W,
                       const uint8_t* dof_shared, bool split_parts, bool relabel_axes, StreamPlan& plan);
void verify_stream_plan(const StreamPlan& plan, const int32_t* tdm, const uint8_t* dof_shared = nullptr);

// Checks every invariant the kernels rely on; throws wfx::Error on violation.
void verify_brick_plan(const BrickPlan& plan, const int32_t* tdm, const uint8_t* dof_shared = nullptr);
void verify_cell_colour_plan(const CellColourPlan& plan, int nd, int64_t ncells, int64_t ndofs,
                             const int32_t* tdm);
} // namespace wfx
