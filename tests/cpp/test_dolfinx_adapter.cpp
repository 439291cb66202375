// Compile-and-run check of the DOLFINx adapter (wave-fenics_b200/hpp/wavefx_dolfinx.hpp) against the
// stand-in classes of mock_dolfinx.hpp: the views must expose exactly the arrays the reference reads
// (common/operators.hpp:53-57, common/precomputation.hpp:29-32, common/cuda/mass.hpp:51), the tagged
// facets must come out as (cell, local facet, tag) and the halo index data as VectorUpdater uses it.
// Needs no GPU.  Prints "adapter ok".
#include "mock_dolfinx.hpp"
#include "wavefx_dolfinx.hpp"

#include <cstdio>
#include <set>
#include <tuple>

#define REQUIRE(cond)                                                              \
  do {                                                                             \
    if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } \
  } while (0)

int main()
{
  const int N = 3, P = 2, n = P + 1, nd = n * n * n, M = P * N + 1;
  auto mesh = std::make_shared<mock::Mesh>();
  std::vector<std::int32_t> xd, xo{0}, dm, dmo{0};
  for (int i = 0; i <= N; ++i)
    for (int j = 0; j <= N; ++j)
      for (int k = 0; k <= N; ++k)
        for (double v : {0.1 * i, 0.1 * j, 0.1 * k}) mesh->_g._x.push_back(v);
  // facet ids: axis-major, then plane, then the two remaining cell coordinates
  auto fid = [&](int axis, int plane, int o1, int o2) { return (axis * (N + 1) + plane) * N * N + o1 * N + o2; };
  const int nfacets = 3 * (N + 1) * N * N;
  std::vector<std::vector<std::int32_t>> f2c(nfacets);
  std::vector<std::int32_t> c2f, c2fo{0};
  for (int cx = 0; cx < N; ++cx)
    for (int cy = 0; cy < N; ++cy)
      for (int cz = 0; cz < N; ++cz)
      {
        const int c = (cx * N + cy) * N + cz;
        for (int v = 0; v < 8; ++v)
          xd.push_back(((cx + (v & 1)) * (N + 1) + (cy + ((v >> 1) & 1))) * (N + 1) + (cz + ((v >> 2) & 1)));
        xo.push_back((std::int32_t)xd.size());
        for (int t = 0; t < nd; ++t) dm.push_back((c * 7 + t) % (M * M * M)); // arbitrary but recognisable
        dmo.push_back((std::int32_t)dm.size());
        // hexahedron facets: 0 z-, 1 y-, 2 x-, 3 x+, 4 y+, 5 z+
        const int fs[6] = {fid(2, cz, cx, cy), fid(1, cy, cx, cz), fid(0, cx, cy, cz),
                           fid(0, cx + 1, cy, cz), fid(1, cy + 1, cx, cz), fid(2, cz + 1, cx, cy)};
        for (int f : fs)
        {
          c2f.push_back(f);
          f2c[f].push_back(c);
        }
        c2fo.push_back((std::int32_t)c2f.size());
      }
  mesh->_g._dm = mock::AdjacencyList(xd, xo);
  std::vector<std::int32_t> f2ca, f2co{0};
  for (auto& l : f2c)
  {
    f2ca.insert(f2ca.end(), l.begin(), l.end());
    f2co.push_back((std::int32_t)f2ca.size());
  }
  mesh->_t.maps[3] = std::make_shared<mock::IndexMap>(N * N * N, 0);
  mesh->_t.conn[{2, 3}] = std::make_shared<mock::AdjacencyList>(f2ca, f2co);
  mesh->_t.conn[{3, 2}] = std::make_shared<mock::AdjacencyList>(c2f, c2fo);

  mock::FunctionSpace V;
  V._mesh = mesh;
  auto dofmap = std::make_shared<mock::DofMap>();
  dofmap->_list = mock::AdjacencyList(dm, dmo);
  auto imap = std::make_shared<mock::IndexMap>(M * M * M - 5, 5);
  dofmap->index_map = imap;
  V._dofmap = dofmap;

  namespace ad = wavefx::dolfinx_adapter;
  wavefx::SpaceView view = ad::make_space_view(V, P);
  REQUIRE(view.degree == P && view.ncells == N * N * N);
  REQUIRE(view.npoints == (N + 1) * (N + 1) * (N + 1));
  REQUIRE(view.x == mesh->geometry().x().data());
  REQUIRE(view.xdofs == mesh->geometry().dofmap().array().data());
  REQUIRE(view.dofmap == dofmap->list().array().data());
  REQUIRE(view.size_local == M * M * M - 5 && view.ndofs == M * M * M);
  REQUIRE(view.nfacets == 0);

  // tags: x = 0 -> 1, x = L -> 2, plus one interior facet that must be skipped
  mock::MeshTags tags;
  std::set<std::tuple<int, int, int>> want;
  for (int cy = 0; cy < N; ++cy)
    for (int cz = 0; cz < N; ++cz)
    {
      tags._i.push_back(fid(0, 0, cy, cz));
      tags._v.push_back(1);
      want.insert({(0 * N + cy) * N + cz, 2, 1});
      tags._i.push_back(fid(0, N, cy, cz));
      tags._v.push_back(2);
      want.insert({((N - 1) * N + cy) * N + cz, 3, 2});
    }
  tags._i.push_back(fid(1, 1, 0, 0));
  tags._v.push_back(7);
  ad::FacetList fl = ad::tagged_facets(*mesh, tags);
  REQUIRE(fl.cell.size() == want.size());
  std::set<std::tuple<int, int, int>> got;
  for (std::size_t i = 0; i < fl.cell.size(); ++i) got.insert({fl.cell[i], fl.local[i], fl.tag[i]});
  REQUIRE(got == want);
  fl.attach(view);
  REQUIRE(view.nfacets == (std::int64_t)want.size() && view.facet_cell == fl.cell.data());

  // halo: 10 owned + 3 ghosts, two destinations, two sources
  mock::IndexMap hm(10, 3);
  hm.fwd = mock::AdjacencyList({4, 7, 9}, {0, 2, 3});
  hm.roff = {0, 1, 3};
  hm.gpos = {2, 0, 1}; // ghost i takes receive-buffer entry gpos[i]
  wavefx::HaloSpec spec = ad::make_halo_spec(hm, {1, 2}, {2, 3});
  REQUIRE((spec.send_offsets == std::vector<std::int32_t>{0, 2, 3}));
  REQUIRE((spec.send_indices == std::vector<std::int32_t>{4, 7, 9}));
  REQUIRE((spec.recv_offsets == std::vector<std::int32_t>{0, 1, 3}));
  REQUIRE((spec.recv_indices == std::vector<std::int32_t>{11, 12, 10}));
  bool threw = false;
  try
  {
    ad::make_halo_spec(hm, {1}, {2, 3});
  }
  catch (const std::runtime_error&)
  {
    threw = true;
  }
  REQUIRE(threw);
  std::printf("adapter ok\n");
  return 0;
}
