// Minimal stand-in for the DOLFINx classes the adapter (wavefx_dolfinx.hpp) touches, with the
// member functions of the API vintage the reference uses.  Test infrastructure only: it lets the
// adapter be compiled and exercised without DOLFINx.  Holds a structured box mesh of hexahedra.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <utility>
#include <vector>

namespace mock
{
template <typename T>
struct Span
{
  const T* p;
  std::size_t n;
  const T* begin() const { return p; }
  const T* end() const { return p + n; }
  std::size_t size() const { return n; }
  const T& operator[](std::size_t i) const { return p[i]; }
};

class AdjacencyList
{
public:
  AdjacencyList() = default;
  AdjacencyList(std::vector<std::int32_t> a, std::vector<std::int32_t> o) : _a(std::move(a)), _o(std::move(o)) {}
  const std::vector<std::int32_t>& array() const { return _a; }
  const std::vector<std::int32_t>& offsets() const { return _o; }
  Span<std::int32_t> links(int i) const { return {_a.data() + _o[i], (std::size_t)(_o[i + 1] - _o[i])}; }

private:
  std::vector<std::int32_t> _a, _o;
};

class IndexMap
{
public:
  IndexMap(std::int32_t nlocal, std::int32_t nghost) : _nl(nlocal), _ng(nghost) {}
  std::int32_t size_local() const { return _nl; }
  std::int32_t num_ghosts() const { return _ng; }
  const AdjacencyList& scatter_fwd_indices() const { return fwd; }
  const std::vector<std::int32_t>& scatter_fwd_receive_offsets() const { return roff; }
  const std::vector<std::int32_t>& scatter_fwd_ghost_positions() const { return gpos; }
  AdjacencyList fwd{{}, {0}};
  std::vector<std::int32_t> roff{0}, gpos;

private:
  std::int32_t _nl, _ng;
};

class Topology
{
public:
  int dim() const { return 3; }
  std::shared_ptr<const IndexMap> index_map(int d) const { return maps.at(d); }
  std::shared_ptr<const AdjacencyList> connectivity(int d0, int d1) const
  {
    auto it = conn.find({d0, d1});
    return it == conn.end() ? nullptr : it->second;
  }
  std::map<int, std::shared_ptr<const IndexMap>> maps;
  std::map<std::pair<int, int>, std::shared_ptr<const AdjacencyList>> conn;
};

class Geometry
{
public:
  const std::vector<double>& x() const { return _x; }
  const AdjacencyList& dofmap() const { return _dm; }
  std::vector<double> _x;
  AdjacencyList _dm;
};

class Mesh
{
public:
  const Topology& topology() const { return _t; }
  const Geometry& geometry() const { return _g; }
  Topology _t;
  Geometry _g;
};

class DofMap
{
public:
  const AdjacencyList& list() const { return _list; }
  std::shared_ptr<const IndexMap> index_map;
  AdjacencyList _list;
};

class FunctionSpace
{
public:
  std::shared_ptr<const Mesh> mesh() const { return _mesh; }
  std::shared_ptr<const DofMap> dofmap() const { return _dofmap; }
  std::shared_ptr<Mesh> _mesh;
  std::shared_ptr<DofMap> _dofmap;
};

class MeshTags
{
public:
  const std::vector<std::int32_t>& indices() const { return _i; }
  const std::vector<std::int32_t>& values() const { return _v; }
  std::vector<std::int32_t> _i, _v;
};
} // namespace mock
