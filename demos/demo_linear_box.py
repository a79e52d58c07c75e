"""Linear wave in a box - the B200 twin of /root/reference/cuda/demo_linear_box.py
(source on x=0, absorbing x=L, degree 4, 80^3 cells, CFL 0.65)."""

import numpy as np

import _common

from fenicsx_fus_gpu_b200 import problem, substrate as S


def main():
    a = _common.parser(__doc__, degree=4, cells=80).parse_args()
    rank, world = _common.init()
    dtype = np.float64 if a.dtype == "f64" else np.float32
    f0, p0, c0, rho = 0.5e6, 60000.0, 1500.0, 1000.0  # :56-63
    L = 0.12  # :66
    h = L / int(2 * L / (c0 / f0))  # the demo's cell size (:85-87)
    grid = S.block_grid(world)
    ncells = tuple(a.cells * g for g in grid)
    lengths = tuple(h * n for n in ncells)
    su = problem.box_setup(a.degree, ncells, lengths, dtype, rank, world, grid=grid)
    solver = problem.linear_solver(su, source_facets=[2], absorbing_facets=[3], rho=rho, c0=c0, f0=f0, p0=p0,
                                   geometry=a.geometry, integrator=a.integrator)
    dt = problem.cfl_time_step(a.degree, h, c0, f0, 0.65)  # :116-120
    tf = lengths[0] / c0 + 2.0 / f0  # :121
    nsteps = a.steps or int(tf / dt) + 1
    if rank == 0:
        print(f"Number of steps: {nsteps}; {su.global_dofs} dofs on {world} GPU(s)", flush=True)
    _common.run(solver, 0.0, dt, nsteps, rank)
    _common.finish(world)


if __name__ == "__main__":
    main()
