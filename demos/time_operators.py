"""Time the operators - the B200 twin of /root/reference/cuda/time_operators.py (CPU twin
numba-cpu/time_operators.py; BASELINE.json configs[0]): mass, boundary-facet mass and
stiffness actions on the unit cube, N^3 hexahedra of degree P, 10 timed launches each
bracketed by a device synchronisation and timed with perf_counter_ns as the reference does
(:205-214, :268-282), same "Elapsed time (...)" lines.  (The CPU side of this config is bench.py's ``cpu_baseline`` /
``--impl reference`` arm.)"""

import argparse
import os
import sys
from time import perf_counter_ns

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--P", type=int, default=4)
    ap.add_argument("--N", type=int, default=32, help="cells per direction (the reference default: 32)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    a = ap.parse_args()
    import torch

    from fenicsx_fus_gpu_b200 import precompute as pre, substrate as S
    from fenicsx_fus_gpu_b200.operators import mass_operator, stiffness_operator

    if torch.cuda.is_available():
        print("CUDA is available")
    print(torch.cuda.get_device_name(0))
    float_type = np.float64 if a.dtype == "f64" else np.float32
    P, N = a.P, a.N
    tb = S.element_tables(P, "basix", float_type)
    mesh = S.create_box(N, 1.0, dtype=float_type)
    dofmap = S.tensor_dofmap(mesh, P)
    num_cells, nd = mesh.num_cells, P + 1
    ndofs = S.num_dofs(N, P)
    print(f"Number of degrees-of-freedom: {ndofs}")
    d = lambda v: torch.from_numpy(np.ascontiguousarray(v)).cuda()  # noqa: E731
    x_dofs_d, x_g_d = d(mesh.x_dofs), d(mesh.x_g)
    tdt = torch.float64 if a.dtype == "f64" else torch.float32
    detJ_d = torch.empty((num_cells, nd**3), dtype=tdt, device="cuda")
    G_d = torch.empty((num_cells, nd**3, 6), dtype=tdt, device="cuda")
    pre.compute_scaled_jacobian_determinant(detJ_d, (x_dofs_d, x_g_d), num_cells, d(tb.dphi), d(tb.wts))
    pre.compute_scaled_geometrical_factor(G_d, (x_dofs_d, x_g_d), num_cells, d(tb.dphi), d(tb.wts))
    boundary_data = np.concatenate([S.boundary_facets(mesh, f) for f in range(6)])
    detJ_f_d = torch.empty((boundary_data.shape[0], nd**2), dtype=tdt, device="cuda")
    pre.compute_boundary_facets_scaled_jacobian_determinant(detJ_f_d, (x_dofs_d, x_g_d), d(boundary_data),
                                                            d(tb.dphi_f), d(tb.wts_f))
    bfacet_dofmap = S.facet_dofmap(dofmap, boundary_data, tb.local_facet_dof)
    dofmap_d, bfacet_dofmap_d = d(dofmap), d(bfacet_dofmap)
    cell_constants_d = torch.ones(num_cells, dtype=tdt, device="cuda")
    bfacet_constants_d = torch.ones(bfacet_dofmap.shape[0], dtype=tdt, device="cuda")
    dphi_1D_d = d(tb.dphi_1D)
    print("Running operators!", flush=True)

    def time10(launch, b_d):
        launch()
        torch.cuda.synchronize()
        t = np.empty(10)
        for i in range(10):
            b_d.zero_()
            torch.cuda.synchronize()
            tic = perf_counter_ns()
            launch()
            torch.cuda.synchronize()
            t[i] = perf_counter_ns() - tic
        return t * 1e-9

    b_d = torch.zeros(ndofs, dtype=tdt, device="cuda")
    u_d = torch.ones(ndofs, dtype=tdt, device="cuda")
    nb = (dofmap.size + 127) // 128
    t_m = time10(lambda: mass_operator[nb, 128](u_d, cell_constants_d, b_d, detJ_d, dofmap_d), b_d)
    print(f"Elapsed time (mass operator): {t_m.mean():.7f} ± {t_m.std():.7f} s")
    nbf = (bfacet_dofmap.size + 127) // 128
    t_f = time10(lambda: mass_operator[nbf, 128](u_d, bfacet_constants_d, b_d, detJ_f_d, bfacet_dofmap_d), b_d)
    print(f"Elapsed time (boundary mass operator): {t_f.mean():.7f} ± {t_f.std():.7f} s")
    xd = S.dof_coordinates(mesh, dofmap, tb)
    u = (100 * np.sin(2 * np.pi * xd[:, 0]) * np.cos(3 * np.pi * xd[:, 1]) * np.sin(4 * np.pi * xd[:, 2])).astype(float_type)
    u_d = d(u)
    stiffness = stiffness_operator(P, float_type)
    t_s = time10(lambda: stiffness[num_cells, (nd, nd, nd)](u_d, cell_constants_d, b_d, G_d, dofmap_d, dphi_1D_d), b_d)
    print(f"Elapsed time (stiffness operator): {t_s.mean():.7f} ± {t_s.std():.7f} s")
    print(f"GDoF/s: mass {ndofs / t_m.mean() / 1e9:.2f}, stiffness {ndofs / t_s.mean() / 1e9:.2f} "
          f"(host-timed, launch + synchronise included)")


if __name__ == "__main__":
    main()
