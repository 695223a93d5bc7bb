"""Copies the reference's lethe-particles application cases that run on a box mesh — parameter file
(input data) plus the final `id, type, dp, x, y, z` table of the golden .output (numbers) — into
tests/golden/apps/ (run where /root/reference is mounted).

    python tests/golden/make_app_goldens.py
"""
import json
import os
import re
import shutil

REF = "/root/reference/applications_tests/lethe-particles"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "apps")
CASES = {
    "rolling_on_plane": "rolling_on_plane.output",
    "velocity_verlet_free_fall": "velocity_verlet_free_fall.output",
    "multiperiodic_collisions_3d": "multiperiodic_collisions_3d.output",
    "multiperiodic_edge_contact_3d": "multiperiodic_edge_contact_3d.output",
    "pp_jkr_equilibrium": "pp_jkr_equilibrium.output",
    "pp_dmt_equilibrium": "pp_dmt_equilibrium.output",
    "pw_jkr_equilibrium": "pw_jkr_equilibrium.output",
    "pw_dmt_equilibrium": "pw_dmt_equilibrium.output",
    "epsd_rolling_resistance_model": "epsd_rolling_resistance_model.output",
    "sliding_in_box": "sliding_in_box.output",
    "periodic_boundary_box": "periodic_boundary_box.output",
    "moving_solid_surface_hmlo": "moving_solid_surface_hmlo.mpirun=1.output",
    "moving_solid_surface_jkr": "moving_solid_surface_jkr.mpirun=1.output",
    "moving_solid_surface_dmt": "moving_solid_surface_dmt.mpirun=1.output",
    "insert_file_3d": "insert_file_3d.mpirun=1.output",
    "insert_list_3d": "insert_list_3d.output",
    "insert_z-x-y": "insert_z-x-y.output",
    "multiperiodic_single_axis_collisions_3d": "multiperiodic_single_axis_collisions_3d.output",
    "single-time-step-list-insertion": "single-time-step-list-insertion.output",
    "periodic_boundary_collisions": "periodic_boundary_collisions.mpirun=1.output",  # == the mpirun=2 golden
    "distribution_normal": "distribution_normal.output",
    "distribution_lognormal": "distribution_lognormal.output",
    "solid_surface": "solid_surface.output",
    "deprecated_parameters": "deprecated_parameters.output",
    "insert_list_3d_default_velocities": "insert_list_3d_default_velocities.output",
    "insertion_acceptance_function": "insertion_acceptance_function.output",
    "insert_plane_3d": "insert_plane_3d.output",  # == the mpirun=2 golden
    "initial_value_insertion": "initial_value_insertion.output",
    "insert_and_remove_with_files": "insert_and_remove_with_files.with_dealii.geq.9.6.mpirun=1.output",
    # a 2-rank golden: the run needs DEMSolver(reference_insertion_ranks=2) for the same site / offset pairing
    "periodic_boundary_load_balancing": "periodic_boundary_load_balancing.mpirun=2.output",
}


def main():
    os.makedirs(OUT, exist_ok=True)
    gold = {}
    for case, output in CASES.items():
        shutil.copy(f"{REF}/{case}.prm", f"{OUT}/{case}.prm")
        rows = []
        for line in open(f"{REF}/{output}"):
            parts = line.split()
            if len(parts) == 6 and parts[0].isdigit():
                rows.append([int(parts[0]), int(parts[1])] + [float(v) for v in parts[2:]])
        gold[case] = rows
        print(case, len(rows), "rows")
    # the moving_solid_surface cases name their mesh `../square.msh` relative to the run directory
    shutil.copy(f"{REF}/moving_solid_surface_files/square.msh", os.path.join(os.path.dirname(OUT), "square.msh"))
    shutil.copy(f"{REF}/insert_file_3d_files/particles.input", os.path.join(os.path.dirname(OUT), "particles.input"))
    for k in (0, 1):
        shutil.copy(f"{REF}/insert_and_remove_with_files_files/particles_0{k}.input", os.path.join(os.path.dirname(OUT), f"particles_0{k}.input"))
    with open(f"{OUT}/final_positions.json", "w") as f:
        json.dump(gold, f)


if __name__ == "__main__":
    main()
