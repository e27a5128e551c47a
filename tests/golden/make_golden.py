"""Generates the golden fixtures in this directory.

    python tests/golden/make_golden.py

The reference (F#, GUI, MathNet) cannot run in this image and ships no fixture files, so the
fixtures pin two things instead:

* ``reference_known_answers.json`` -- values that follow from the reference's own source by
  hand: its three still-valid NUnit known-answer tests (Tests.fs:54-73), the `Coord.nearest`
  spot values (Coord.fs:49-57) and the stock N = 1 `advance` step.  These are
  typed in below, NOT computed by the oracle; the oracle is checked against them.
* ``oracle_<scene>.npz`` -- outputs of the CPU oracle (`oracle/dense_ref.py`, fp64) on small
  seeded scenes: advected field, right-hand side, directly solved pressure, projected
  field, one multigrid V-cycle, marker stage.  They freeze the oracle so that a later edit of
  the restatement cannot silently move the goal posts; the GPU parity tests check the CUDA
  path against these files as well as against the live oracle.
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

from oracle import dense_ref as dr  # noqa: E402
from splashy_b200 import scenes  # noqa: E402

KNOWN = {
    "source": "hand-derived from /root/reference/src/splashy (SURVEY.md App. B, Tests.fs:54-73)",
    "nunit_divergence": [
        {"cite": "Tests.fs:54-57", "neighbours": "Air", "velocity": [0.0, 0.0, 0.0], "divergence": 0.0},
        {"cite": "Tests.fs:59-64", "neighbours": "Air", "velocity": [1.0, 2.0, 3.0],
         "divergence": -6.0, "note": "in - out with only the own +faces set: -(1 + 2 + 3)"},
        {"cite": "Tests.fs:66-73", "neighbours": "Solid", "velocity": [1.0, 2.0, 3.0],
         "divergence": 0.0, "note": "every face is gated by a solid neighbour"},
    ],
    "coord_nearest": {"3": 0, "4": 0, "5": 10, "14": 10, "15": 20, "-3": 0, "-5": 0, "-7": 0,
                      "-10": -10, "-11": -10, "-15": -10, "-19": -10},
    "construct_float": {"input": [4.5, 5.5, -5.5], "coord": [0, 10, 0]},
    "stock_step": {
        "dt": 0.016671,
        "after_forces": [0.0, -0.16354251, 0.0],
        "after_viscosity_y": -0.1635424936414969,
        "h_rho_over_dt": 599826.0452282408,
        "divergence": 0.1635424936414969,
        "A": [[6.0]],
        "b": 98097.0471877438,
        "pressure": 16349.507864623965,
        "fluid_velocity_after_apply": [0.027257082273582815, -0.13628541136791408,
                                       0.027257082273582815],
        "neighbour_own_axis_after_apply": -0.027257082273582815,
    },
}


def oracle_fixture(name, sc, markers=8, seed=0):
    lab = sc["labels"]
    U, V, W = (sc[k].astype(np.float64) for k in "UVW")
    dt, h, rho = sc["dt"], sc["h"], sc["rho"]
    U1, V1, W1 = dr.advect(lab, U, V, W, dt, h)
    b = dr.rhs(lab, U1, V1, W1, dt, h, rho)
    P = dr.solve_direct(lab, b)
    U2, V2, W2 = dr.grad_sub(lab, U1, V1, W1, P, dt, h, rho)
    rng = np.random.default_rng(seed)
    r = np.where(lab == dr.FLUID, rng.uniform(-1, 1, lab.shape), 0.0)
    z = dr.Multigrid(lab, nu=2).vcycle(r)
    nz, ny, nx = lab.shape
    inner = np.zeros(lab.shape, dtype=bool)              # markers start off the box faces and
    inner[min(1, nz - 1):max(nz - 1, 1), 1:ny - 1, 1:nx - 1] = True   # move < 1 cell: they stay inside
    fl = np.argwhere((lab == dr.FLUID) & inner)
    pick = fl[rng.choice(len(fl), size=min(markers, len(fl)), replace=False)]
    cells = pick[:, ::-1]                                   # (i, j, k)
    loc = cells * float(h) + rng.uniform(-4.0, 4.0, cells.shape)
    vmax = max(np.abs(U).max(), np.abs(V).max(), np.abs(W).max(), 1e-30)
    mdt = 0.9 * float(h) / vmax
    new_loc, old, new = dr.move_markers(lab, U, V, W, loc, mdt, h, (0, 0, 0))
    assert (new >= 0).all() and (new < np.array([nx, ny, nz])).all() and (old != new).any()
    lab2, Um, Vm, Wm, Pm = dr.relabel(lab, U, V, W, P * 0, new, (0, 0, 0), (nx - 1, ny - 1, nz - 1))
    np.savez_compressed(HERE / f"oracle_{name}.npz", labels=lab, U=U, V=V, W=W, dt=dt, h=h, rho=rho,
                        U_adv=U1, V_adv=V1, W_adv=W1, rhs=b, P=P, U_proj=U2, V_proj=V2, W_proj=W2,
                        mg_r=r, mg_z=z, marker_loc=loc, marker_dt=mdt, marker_new_loc=new_loc,
                        relabel_labels=lab2)


def main():
    (HERE / "reference_known_answers.json").write_text(json.dumps(KNOWN, indent=1) + "\n")
    oracle_fixture("vortex_16x12x10", scenes.vortex(16, 12, 10, cfl=1.5, dtype=np.float64))
    oracle_fixture("dambreak_12", scenes.dambreak(12, dtype=np.float64))
    oracle_fixture("smoke2d_24", scenes.smoke2d(24, dtype=np.float64))
    oracle_fixture("random_13x9x7", scenes.random_box(13, 9, 7, seed=3, dtype=np.float64,
                                                      p_solid=0.15, p_air=0.3))


if __name__ == "__main__":
    main()
