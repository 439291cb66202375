// DOLFINx adapter: builds the plain-array views of wavefx.hpp from the
// DOLFINx objects the reference passes around, using only the member functions the reference itself
// calls (API vintage ~ v0.3.1.dev / v0.4.0, see the citations).  The functions are templates on the
// DOLFINx types, so this header has no DOLFINx include of its own: it compiles against the real
// library and against the minimal stand-in in tests/cpp/mock_dolfinx.hpp, which is how it is
// compile-checked in this repository (DOLFINx itself is not available here; the adapter has NOT
// been run against a real DOLFINx build).
//
//   auto view  = wavefx::dolfinx_adapter::make_space_view(*V, degree);           // operators.hpp:53-57,149-153
//   auto fl    = wavefx::dolfinx_adapter::tagged_facets(*V->mesh(), *meshtags);  // LinearGLL.hpp:113-115
//   fl.attach(view);
//   auto spec  = wavefx::dolfinx_adapter::make_halo_spec(*V->dofmap()->index_map, send_ranks, recv_ranks);
#pragma once

#include "wavefx.hpp"

#include <algorithm>
#include <stdexcept>

namespace wavefx
{
namespace dolfinx_adapter
{
// Geometry, dofmap and sizes of a function space (facets left empty).
template <class FunctionSpace>
SpaceView make_space_view(const FunctionSpace& V, int degree)
{
  auto mesh = V.mesh();
  const int tdim = mesh->topology().dim();
  SpaceView v;
  v.degree = degree;
  v.ncells = mesh->topology().index_map(tdim)->size_local();              // operators.hpp:57
  const auto& x = mesh->geometry().x();                                    // precomputation.hpp:29-31
  v.npoints = (std::int64_t)(x.size() / 3);
  v.x = x.data();
  v.xdofs = mesh->geometry().dofmap().array().data();                      // precomputation.hpp:32
  auto map = V.dofmap()->index_map;
  v.size_local = map->size_local();
  v.ndofs = v.size_local + map->num_ghosts();                              // LinearGLL.hpp:81,106
  v.dofmap = V.dofmap()->list().array().data();                            // cuda/mass.hpp:51
  return v;
}

// (cell, local facet, tag) triplets of the tagged exterior facets: what the FFCx facet kernel of the
// reference iterates over (demo/cpu_planar3d/forms.ufl:21-24, LinearGLL.hpp:113-115,175).
struct FacetList
{
  std::vector<std::int32_t> cell, local, tag;
  void attach(SpaceView& v) const
  {
    v.nfacets = (std::int64_t)cell.size();
    v.facet_cell = cell.data();
    v.facet_local = local.data();
    v.facet_tag = tag.data();
  }
};

template <class Mesh, class MeshTags>
FacetList tagged_facets(const Mesh& mesh, const MeshTags& tags)
{
  const int tdim = mesh.topology().dim();
  auto f_to_c = mesh.topology().connectivity(tdim - 1, tdim);
  auto c_to_f = mesh.topology().connectivity(tdim, tdim - 1);
  if (!f_to_c || !c_to_f) throw std::runtime_error("wavefx: facet-cell connectivity has not been created");
  const std::int32_t ncells_local = mesh.topology().index_map(tdim)->size_local();
  FacetList out;
  const auto& facets = tags.indices();
  const auto& values = tags.values();
  for (std::size_t i = 0; i < facets.size(); ++i)
  {
    auto cells = f_to_c->links(facets[i]);
    if (cells.size() != 1) continue; // interior facet
    const std::int32_t c = cells[0];
    if (c >= ncells_local) continue; // ghost cell: its owner integrates the facet
    auto cf = c_to_f->links(c);
    auto it = std::find(cf.begin(), cf.end(), facets[i]);
    if (it == cf.end()) throw std::runtime_error("wavefx: inconsistent facet connectivity");
    out.cell.push_back(c);
    out.local.push_back((std::int32_t)(it - cf.begin()));
    out.tag.push_back((std::int32_t)values[i]);
  }
  return out;
}

// The index data VectorUpdater's constructor reads (demo/gpu_scatter_mpi/VectorUpdater.hpp:31-59).
// send_ranks / recv_ranks: destinations / sources of the forward neighbourhood communicator, in its
// order (MPI_Dist_graph_neighbors on index_map.comm(Direction::forward), :69-80 of the same file).
template <class IndexMap>
HaloSpec make_halo_spec(const IndexMap& map, const std::vector<std::int32_t>& send_ranks,
                        const std::vector<std::int32_t>& recv_ranks)
{
  HaloSpec s;
  s.send_ranks = send_ranks;
  s.recv_ranks = recv_ranks;
  const auto& shared = map.scatter_fwd_indices();
  s.send_offsets.assign(shared.offsets().begin(), shared.offsets().end());
  s.send_indices.assign(shared.array().begin(), shared.array().end());
  const auto& roff = map.scatter_fwd_receive_offsets();
  s.recv_offsets.assign(roff.begin(), roff.end());
  const auto& gpos = map.scatter_fwd_ghost_positions();
  const std::int32_t size_local = map.size_local();
  s.size_local = size_local;
  s.num_ghosts = map.num_ghosts();
  // ghost i takes entry gpos[i] of the receive buffer (update_fwd_end gathers with these positions,
  // VectorUpdater.hpp:138-142, scatter.cu:5-10); wfx_halo_create wants the inverse: the local slot
  // filled by each receive-buffer entry
  s.recv_indices.assign(gpos.size(), -1);
  for (std::size_t i = 0; i < gpos.size(); ++i)
  {
    if (gpos[i] < 0 || (std::size_t)gpos[i] >= gpos.size() || s.recv_indices[gpos[i]] >= 0)
      throw std::runtime_error("wavefx: scatter_fwd_ghost_positions is not a permutation");
    s.recv_indices[gpos[i]] = size_local + (std::int32_t)i;
  }
  if (s.send_offsets.size() != s.send_ranks.size() + 1 || s.recv_offsets.size() != s.recv_ranks.size() + 1)
    throw std::runtime_error("wavefx: neighbour rank lists do not match the index map's offsets");
  return s;
}
} // namespace dolfinx_adapter
} // namespace wavefx
