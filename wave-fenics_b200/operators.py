"""Host-side mirror of the reference operator interface over the C ABI.

Same names, argument meaning and call shape as the reference classes:

  MassOperator(V, degree)(x, y)                 common/operators.hpp:43-109 (MassOperatorCPU)
  StiffnessOperator(V, degree, params)(x, y)    common/operators.hpp:136-201
  LinearGLLOpt(mesh, meshtags, degree, c0, f0, p0).init() / .rk4(t0, tf, dt)
                                                common/LinearGLL.hpp:37-287

`V` is a HexMesh (mesh.py): it carries what the reference pulls out of the DOLFINx
FunctionSpace (geometry, geometry dofmap, cell dofmap, index-map sizes).  Vectors are
DOLFINx-layout arrays: torch CUDA tensors (device path, asynchronous on the current torch
stream) or numpy arrays (host path: copied to the GPU and back inside the call, like the
reference functor applied to host la::Vectors).  All compute happens in libwavefx.so.
"""
import ctypes as C

import numpy as np

from . import capi


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Context:
    """One per GPU (role of utils::set_device, common/cuda/utils.hpp:22-38)."""
    _cache = {}

    def __init__(self, device=0):
        self.device = int(device)
        self.handle = C.c_void_p()
        capi.call("wfx_ctx_create", self.device, C.byref(self.handle))

    @classmethod
    def get(cls, device=0):
        device = int(device)
        if device not in cls._cache:
            cls._cache[device] = cls(device)
        return cls._cache[device]

    def synchronize(self):
        capi.call("wfx_ctx_sync", self.handle)


class Geometry:
    """precompute_geometric_data (common/precomputation.hpp:18-110) on the GPU."""

    def __init__(self, V, degree, dtype=np.float64, ctx=None):
        self.ctx = ctx or Context.get()
        self.P = int(degree)
        self.dtype = np.dtype(dtype)
        self.ncells = V.ncells
        self.handle = C.c_void_p()
        x = np.ascontiguousarray(V.x, dtype=np.float64)
        xd = np.ascontiguousarray(V.xdofs, dtype=np.int32)
        capi.call("wfx_geometry_create", self.ctx.handle, self.P, capi.dtype_code(self.dtype),
                  V.ncells, x.shape[0], capi.f64p(x.reshape(-1)), capi.i32p(xd.reshape(-1)),
                  C.byref(self.handle))

    def info(self):
        nc, na = C.c_int64(), C.c_int64()
        capi.call("wfx_geometry_info", self.handle, C.byref(nc), C.byref(na))
        return dict(ncells=nc.value, n_affine=na.value)

    def scale_cells(self, coeff):
        """G[c] *= coeff[c]: a piecewise-constant coefficient, e.g. (c0[c] / c0_ref)^2."""
        coeff = np.ascontiguousarray(coeff, dtype=np.float64)
        if coeff.size != self.ncells:
            raise capi.WfxError("one coefficient per cell expected")
        capi.call("wfx_geometry_scale_cells", self.handle, capi.f64p(coeff))

    def get(self):
        """(G [ncells,nq,3,3], detJ [ncells,nq]) in the reference layout."""
        nq = (self.P + 1) ** 3
        G = np.empty((self.ncells, nq, 3, 3))
        detJ = np.empty((self.ncells, nq))
        capi.call("wfx_geometry_get", self.handle, capi.f64p(G.reshape(-1)), capi.f64p(detJ.reshape(-1)))
        return G, detJ

    def __del__(self):
        if getattr(self, "handle", None) and getattr(capi, "lib", None) is not None:
            capi.lib.wfx_geometry_destroy(self.handle)
            self.handle = None


def compute_jacobian_data(V, points, weights=None, ctx=None, want=("J", "detJ", "K", "G")):
    """compute_jacobian / _determinant / _inverse / compute_geometrical_factor
    (common/precompute.hpp:49-176) at arbitrary reference points, on the GPU."""
    ctx = ctx or Context.get()
    points = np.ascontiguousarray(points, dtype=np.float64)
    nq = points.shape[0]
    nc = V.ncells
    out = {}
    J = np.empty((nc, nq, 3, 3)) if "J" in want else None
    detJ = np.empty((nc, nq)) if "detJ" in want else None
    K = np.empty((nc, nq, 3, 3)) if "K" in want else None
    G = np.empty((nc, nq, 3, 3)) if "G" in want and weights is not None else None
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    x = np.ascontiguousarray(V.x, dtype=np.float64)
    xd = np.ascontiguousarray(V.xdofs, dtype=np.int32)
    flat = lambda a: None if a is None else capi.f64p(a.reshape(-1))
    capi.call("wfx_compute_jacobian_data", ctx.handle, nc, x.shape[0], capi.f64p(x.reshape(-1)),
              capi.i32p(xd.reshape(-1)), nq, capi.f64p(points.reshape(-1)), flat(w), flat(J),
              flat(detJ), flat(K), flat(G))
    for k, v in (("J", J), ("detJ", detJ), ("K", K), ("G", G)):
        if v is not None:
            out[k] = v
    return out


class _Operator:
    dtype = np.dtype(np.float64)
    ndofs = 0

    def _check(self, x, y):
        if _is_torch(x) != _is_torch(y):
            raise capi.WfxError("x and y must both be torch CUDA tensors or both numpy arrays")
        if _is_torch(x):
            import torch
            want = torch.float64 if self.dtype == np.float64 else torch.float32
            for v in (x, y):
                if not v.is_cuda or v.dtype != want or not v.is_contiguous() or v.numel() != self.ndofs:
                    raise capi.WfxError("vector must be a contiguous CUDA tensor of the operator's dtype and size")
            return True
        for v in (x, y):
            if v.dtype != self.dtype or not v.flags["C_CONTIGUOUS"] or v.size != self.ndofs:
                raise capi.WfxError("vector must be a contiguous numpy array of the operator's dtype and size")
        return False


class StiffnessOperator(_Operator):
    """y += -c0^2 K x  (common/operators.hpp:183-200).  `params` is accepted and ignored for
    the speed of sound exactly as the reference does (c0 = 1500 hard-coded, :113-115) unless
    honour_params=True."""

    def __init__(self, V, degree, params=None, dtype=np.float64, ctx=None, geometry=None,
                 mode=capi.STIFF_AUTO, honour_params=False, split=True):
        self.ctx = ctx or Context.get()
        self.P = int(degree)
        self.dtype = np.dtype(dtype)
        self.geometry = geometry or Geometry(V, degree, dtype, self.ctx)
        self.ndofs = int(V.ndofs)
        c0 = 1500.0
        if honour_params and params and "c0" in params:
            c0 = float(params["c0"])
        self.c0 = c0
        self.handle = C.c_void_p()
        dm = np.ascontiguousarray(V.dofmap, dtype=np.int32)
        # distributed mesh: the dofs that also live on another rank (halo send + receive lists)
        halo = getattr(V, "halo", None) or {}
        shared = None
        if len(halo.get("send_indices", ())) or len(halo.get("recv_indices", ())):
            shared = np.unique(np.concatenate([halo["send_indices"], halo["recv_indices"]])).astype(np.int32)
        self.nshared = 0 if shared is None else len(shared)
        capi.call("wfx_stiffness_create_partitioned", self.ctx.handle, self.geometry.handle, self.ndofs,
                  capi.i32p(dm.reshape(-1)), c0, mode | (0 if split else capi.STIFF_NO_SPLIT), self.nshared,
                  capi.i32p(shared), C.byref(self.handle))
        self.split = bool(split) and self.nshared > 0

    def __call__(self, x, y):
        self.apply(x, y, beta=1)

    def apply(self, x, y, beta=1):
        if self._check(x, y):
            capi.call("wfx_stiffness_apply", self.handle, C.c_void_p(x.data_ptr()),
                      C.c_void_p(y.data_ptr()), int(beta), _stream_ptr())
        else:
            capi.call("wfx_stiffness_apply_host", self.handle, C.c_void_p(x.ctypes.data),
                      C.c_void_p(y.ctypes.data), int(beta))

    def apply_part(self, x, y, part, beta=0, scale_ptr=None, stream=None):
        """Distributed meshes: part 0 = cells touching rank-shared dofs, 1 = the rest, -1 = both.
        scale_ptr (1/m) is applied on the last touch of NON-shared dofs only."""
        self._check(x, y)
        capi.call("wfx_stiffness_apply_part", self.handle, C.c_void_p(x.data_ptr()),
                  C.c_void_p(scale_ptr) if scale_ptr else None, C.c_void_p(y.data_ptr()), int(beta),
                  int(part), stream if stream is not None else _stream_ptr())

    def apply_scaled(self, x, scale_ptr, y):
        """y = scale .* (-c0^2 K x): the fused stiffness + lumped-mass-inverse apply."""
        self._check(x, y)
        capi.call("wfx_stiffness_apply_scaled", self.handle, C.c_void_p(x.data_ptr()),
                  C.c_void_p(scale_ptr), C.c_void_p(y.data_ptr()), _stream_ptr())

    def info(self):
        nc, nd, ndofs = C.c_int64(), C.c_int(), C.c_int64()
        fl, by, ncol, nl = C.c_double(), C.c_double(), C.c_int(), C.c_int()
        capi.call("wfx_stiffness_info", self.handle, C.byref(nc), C.byref(nd), C.byref(ndofs),
                  C.byref(fl), C.byref(by), C.byref(ncol), C.byref(nl))
        return dict(num_cells=nc.value, num_dofs=nd.value, ndofs=ndofs.value, flops=fl.value,
                    bytes=by.value, ncolours=ncol.value, nlaunches=nl.value)

    def kernel_info(self):
        v, a, m, nr, nb, sm = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        capi.call("wfx_stiffness_kernel_info", self.handle, C.byref(v), C.byref(a), C.byref(m), C.byref(nr),
                  C.byref(nb), C.byref(sm))
        names = {-1: "cell", 0: "brick-generic", 1: "brick-regular", 2: "brick-regular-p4d", 3: "cell-streamed", 4: "brick-streamed"}
        return dict(variant=names.get(v.value, str(v.value)), affine=bool(a.value), mixed=bool(m.value),
                    regular_batches=nr.value, batches=nb.value, smem_bytes=sm.value)

    def __del__(self):
        if getattr(self, "handle", None) and getattr(capi, "lib", None) is not None:
            capi.lib.wfx_stiffness_destroy(self.handle)
            self.handle = None


class MassOperator(_Operator):
    """y += M x with the collocated (diagonal) GLL mass (common/operators.hpp:86-108)."""

    def __init__(self, V, degree, dtype=np.float64, ctx=None, geometry=None):
        self.ctx = ctx or Context.get()
        self.P = int(degree)
        self.dtype = np.dtype(dtype)
        self.geometry = geometry or Geometry(V, degree, dtype, self.ctx)
        self.ndofs = int(V.ndofs)
        self.handle = C.c_void_p()
        dm = np.ascontiguousarray(V.dofmap, dtype=np.int32)
        capi.call("wfx_mass_create", self.ctx.handle, self.geometry.handle, self.ndofs,
                  capi.i32p(dm.reshape(-1)), C.byref(self.handle))

    def __call__(self, x, y):
        self.apply(x, y, beta=1)

    def apply(self, x, y, beta=1):
        if self._check(x, y):
            capi.call("wfx_mass_apply", self.handle, C.c_void_p(x.data_ptr()),
                      C.c_void_p(y.data_ptr()), int(beta), _stream_ptr())
        else:
            capi.call("wfx_mass_apply_host", self.handle, C.c_void_p(x.ctypes.data),
                      C.c_void_p(y.ctypes.data), int(beta))

    def apply_inverse(self, x, y):
        """y = x ./ m (x and y may be the same tensor)."""
        self._check(x, y)
        capi.call("wfx_mass_apply_inverse", self.handle, C.c_void_p(x.data_ptr()),
                  C.c_void_p(y.data_ptr()), _stream_ptr())

    def assemble(self, halo):
        """Distributed meshes: sum the diagonal over the ranks sharing a dof (fp64 halo)."""
        capi.call("wfx_mass_assemble", self.handle, halo.handle)

    def _ptr(self, name):
        p = C.c_void_p()
        capi.call(name, self.handle, C.byref(p))
        return p.value

    def diagonal_ptr(self):
        return self._ptr("wfx_mass_diagonal")

    def inverse_diagonal_ptr(self):
        return self._ptr("wfx_mass_inverse_diagonal")

    def _download(self, ptr):
        out = np.empty(self.ndofs, dtype=self.dtype)
        capi.call("wfx_memcpy_d2h", self.ctx.handle, C.c_void_p(out.ctypes.data), C.c_void_p(ptr), out.nbytes)
        return out

    def diagonal(self):
        return self._download(self.diagonal_ptr())

    def inverse_diagonal(self):
        return self._download(self.inverse_diagonal_ptr())

    def __del__(self):
        if getattr(self, "handle", None) and getattr(capi, "lib", None) is not None:
            capi.lib.wfx_mass_destroy(self.handle)
            self.handle = None


# the reference names both spellings (LinearGLL.hpp:63 vs operators.hpp:44)
MassOperatorCPU = MassOperator


class BoundaryOperator(_Operator):
    """b += c0^2 g m1 - c0 m2 .* v_n  (forms.ufl:21-24 via assemble_vector, LinearGLL.hpp:175)."""

    def __init__(self, V, degree, dtype=np.float64, ctx=None):
        self.ctx = ctx or Context.get()
        self.P = int(degree)
        self.dtype = np.dtype(dtype)
        self.ndofs = int(V.ndofs)
        self.handle = C.c_void_p()
        x = np.ascontiguousarray(V.x, dtype=np.float64)
        xd = np.ascontiguousarray(V.xdofs, dtype=np.int32)
        dm = np.ascontiguousarray(V.dofmap, dtype=np.int32)
        fc = np.ascontiguousarray(V.facet_cells, dtype=np.int32)
        fl = np.ascontiguousarray(V.facet_local, dtype=np.int32)
        ft = np.ascontiguousarray(V.facet_tags, dtype=np.int32)
        capi.call("wfx_boundary_create", self.ctx.handle, self.P, capi.dtype_code(self.dtype), len(fc),
                  capi.i32p(fc), capi.i32p(fl), capi.i32p(ft), x.shape[0], capi.f64p(x.reshape(-1)),
                  capi.i32p(xd.reshape(-1)), self.ndofs, capi.i32p(dm.reshape(-1)), C.byref(self.handle))

    def apply(self, c0, g, vn, b):
        self._check(vn, b)
        capi.call("wfx_boundary_apply", self.handle, float(c0), float(g), C.c_void_p(vn.data_ptr()),
                  C.c_void_p(b.data_ptr()), _stream_ptr())

    def assemble(self, halo):
        """Distributed meshes: sum the facet masses over the ranks sharing a boundary dof (fp64 halo)."""
        capi.call("wfx_boundary_assemble", self.handle, halo.handle)

    def facet_masses(self):
        m1, m2 = np.empty(self.ndofs), np.empty(self.ndofs)
        capi.call("wfx_boundary_get", self.handle, capi.f64p(m1), capi.f64p(m2))
        return m1, m2

    def __del__(self):
        if getattr(self, "handle", None) and getattr(capi, "lib", None) is not None:
            capi.lib.wfx_boundary_destroy(self.handle)
            self.handle = None


class LinearGLLOpt:
    """The wave model + RK4 driver (common/LinearGLL.hpp:37-287) on the GPU.

    mesh carries the facet tags (`meshtags` may be None or a (cells, local, tags) triple that
    overrides them)."""

    def __init__(self, mesh, meshtags, degreeOfBasis, speedOfSound, sourceFrequency,
                 pressureAmplitude, dtype=np.float64, ctx=None, halo=None, stiffness_mode=capi.STIFF_AUTO,
                 setup_halo=None):
        self.ctx = ctx or Context.get()
        self.V = mesh
        if meshtags is not None:
            import copy
            mesh = copy.copy(mesh)
            mesh.facet_cells, mesh.facet_local, mesh.facet_tags = meshtags
        self.k_ = int(degreeOfBasis)
        self.c0_, self.freq0_, self.p0_ = float(speedOfSound), float(sourceFrequency), float(pressureAmplitude)
        self.dtype = np.dtype(dtype)
        self.geometry = Geometry(mesh, self.k_, dtype, self.ctx)
        self.mass_op = MassOperator(mesh, self.k_, dtype, self.ctx, self.geometry)      # :105
        params = {"c0": self.c0_}
        self.stiff_op = StiffnessOperator(mesh, self.k_, params, dtype, self.ctx, self.geometry,
                                          mode=stiffness_mode)                           # :120
        self.bnd_op = BoundaryOperator(mesh, self.k_, dtype, self.ctx)                   # :113-115
        self.halo = halo
        if halo is not None:
            # the masses are summed over the ranks in fp64: an fp32 model needs a second, fp64 halo
            # for this one-off set-up step (`setup_halo`, built here on the same mesh if not given)
            if setup_halo is None:
                if self.dtype == np.float64:
                    setup_halo = halo
                else:
                    from .partition import Halo
                    setup_halo = Halo(mesh, self.ctx, np.float64, comm=halo.comm)
            self.setup_halo = setup_halo
            self.mass_op.assemble(setup_halo)                                            # :110
            self.bnd_op.assemble(setup_halo)
        self.handle = C.c_void_p()
        capi.call("wfx_wave_create", self.ctx.handle, self.stiff_op.handle, self.mass_op.handle,
                  self.bnd_op.handle, halo.handle if halo is not None else None, int(mesh.size_local),
                  self.c0_, self.freq0_, self.p0_, C.byref(self.handle))
        self.ndofs = int(mesh.ndofs)

    def init(self):
        capi.call("wfx_wave_init", self.handle)

    def set_state(self, u, v):
        u = np.ascontiguousarray(u, dtype=self.dtype)
        v = np.ascontiguousarray(v, dtype=self.dtype)
        capi.call("wfx_wave_set_state", self.handle, C.c_void_p(u.ctypes.data), C.c_void_p(v.ctypes.data))

    def get_state(self):
        u, v = np.empty(self.ndofs, dtype=self.dtype), np.empty(self.ndofs, dtype=self.dtype)
        capi.call("wfx_wave_get_state", self.handle, C.c_void_p(u.ctypes.data), C.c_void_p(v.ctypes.data))
        return u, v

    def set_probes(self, dofs, max_records):
        """Record u[dofs] after every completed time step (kept on the device, up to max_records)."""
        dofs = np.ascontiguousarray(dofs, dtype=np.int32)
        self._nprobes = len(dofs)
        capi.call("wfx_wave_set_probes", self.handle, len(dofs), capi.i32p(dofs), int(max_records))

    def probe_series(self):
        """(t [nrec], u [nrec, nprobes]) recorded so far."""
        n = C.c_int64()
        capi.call("wfx_wave_get_probe_series", self.handle, C.byref(n), None, None)
        t = np.empty(n.value)
        vals = np.empty((n.value, getattr(self, "_nprobes", 0)), dtype=self.dtype)
        capi.call("wfx_wave_get_probe_series", self.handle, C.byref(n), capi.f64p(t), C.c_void_p(vals.ctypes.data))
        return t, vals

    def set_snapshot(self, every, fn):
        """fn(step, t, u, v) with numpy views of the pinned host copies (copy them to keep them), every
        `every` completed steps; the device -> host copy overlaps the time stepping."""
        if fn is None or not every:
            self._snap_cb = None
            capi.call("wfx_wave_set_snapshot", self.handle, 0, None, None)
            return
        n, dt = self.ndofs, self.dtype

        def tramp(_user, step, t, up, vp):
            ctype = C.c_double if dt == np.float64 else C.c_float
            u = np.ctypeslib.as_array(C.cast(up, C.POINTER(ctype)), shape=(n,))
            v = np.ctypeslib.as_array(C.cast(vp, C.POINTER(ctype)), shape=(n,))
            fn(int(step), float(t), u, v)

        self._snap_cb = capi.SNAPSHOT_FN(tramp)  # keep the trampoline alive
        capi.call("wfx_wave_set_snapshot", self.handle, int(every), C.cast(self._snap_cb, C.c_void_p), None)

    def f0(self, t, u, v, result):
        """result = v (LinearGLL.hpp:141-144); device tensors."""
        capi.call("wfx_wave_f0", self.handle, float(t), C.c_void_p(u.data_ptr()), C.c_void_p(v.data_ptr()),
                  C.c_void_p(result.data_ptr()), _stream_ptr())

    def f1(self, t, u, v, result):
        """result = M^-1 (-c0^2 K u + boundary(g(t), v)) (LinearGLL.hpp:151-192); device tensors."""
        capi.call("wfx_wave_f1", self.handle, float(t), C.c_void_p(u.data_ptr()), C.c_void_p(v.data_ptr()),
                  C.c_void_p(result.data_ptr()), _stream_ptr())

    def rk4(self, startTime, finalTime, timeStep, max_steps=0, stream=None):
        steps, t_end = C.c_int64(), C.c_double()
        capi.call("wfx_wave_rk4", self.handle, float(startTime), float(finalTime), float(timeStep),
                  int(max_steps), C.byref(steps), C.byref(t_end), stream if stream is not None else _stream_ptr())
        return steps.value, t_end.value

    def __del__(self):
        if getattr(self, "handle", None) and getattr(capi, "lib", None) is not None:
            capi.lib.wfx_wave_destroy(self.handle)
            self.handle = None
