"""Extracts the known-answer vectors the reference's own tests hold for the DEM
hot path into small JSON fixtures (run in the build container, where
/root/reference is mounted; the GPU box only sees the committed JSON).

    python tests/golden/make_golden.py

Sources (all under /root/reference):
  tests/dem/*.output                                  unit-test goldens (6 significant digits)
  applications_tests/lethe-particles/*.output         end-state positions (4 decimals)
Only numbers are extracted; no reference source code is copied.
"""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
NUM = r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?"


def read(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read()


def floats(s):
    return [float(v) for v in re.findall(NUM, s)]


def full_contact():
    txt = read("tests/dem/particle_particle_full_contact.output")
    series = {}
    cur = None
    for line in txt.splitlines():
        m = re.search(r"For the (\w+) contact force model", line)
        if m:
            cur = m.group(1)
            series[cur] = []
            continue
        m = re.search(r"At time (\S+)\s+force is (.*) torque is (.*) overlap is (.*)$", line)
        if m and cur:
            series[cur].append(
                {"time": float(m.group(1)), "force": floats(m.group(2)), "torque": floats(m.group(3)), "overlap": float(m.group(4))}
            )
    return series


def application_positions(name):
    txt = read(f"applications_tests/lethe-particles/{name}.output")
    rows = []
    seen_header = False
    for line in txt.splitlines():
        if line.startswith("id, type, dp, x, y, z"):
            seen_header = True
            rows = []
            continue
        if seen_header:
            parts = line.split()
            if len(parts) == 6:
                try:
                    rows.append([int(parts[0]), int(parts[1])] + [float(v) for v in parts[2:]])
                except ValueError:
                    pass
    return rows


def contact_on_two_processors():
    """tests/dem/particle_particle_contact_on_two_processors.mpirun=2.output: y of particle 0 every
    10th step of a head-on collision across the boundary between the two ranks."""
    txt = read("tests/dem/particle_particle_contact_on_two_processors.mpirun=2.output")
    return [float(m) for m in re.findall(r"location of particle 0 is: \S+ (\S+)", txt)]


def main():
    g = {}
    g["pp_force_nonlinear"] = floats(read("tests/dem/particle_particle_contact_force_nonlinear.output").split("is:")[1])[:3]
    g["pp_force_linear"] = floats(read("tests/dem/particle_particle_contact_force_linear.output").split("is:")[1])[:3]
    g["pw_force_nonlinear"] = floats(read("tests/dem/particle_wall_contact_force_nonlinear.output").split("is:")[1])[0]
    g["pw_force_linear"] = floats(read("tests/dem/particle_wall_contact_force_linear.output").split("is:")[1])[0]
    g["full_contact"] = full_contact()
    g["normal_force"] = [float(l.split("::")[1]) for l in read("tests/dem/normal_force.output").splitlines() if "::" in l and l.split("::")[1].strip()]
    pcv = read("tests/dem/post_collision_velocity.output")
    g["post_collision_velocity"] = [
        {"restitution": float(m.group(1)), "v_after": float(m.group(2))}
        for m in re.finditer(r"restitution is (\S+) and .* collision is: (\S+)", pcv)
    ]
    vv = [l for l in read("tests/dem/integration_velocity_verlet.output").splitlines() if l.startswith("DEAL::")]
    start = [i for i, l in enumerate(vv) if "Final time" in l][1]  # second block = dim 3
    blk = vv[start : start + 4]
    g["velocity_verlet_3d"] = {
        "final_time": floats(blk[0].split("time:")[1])[0],
        "position": floats(blk[1].split("Position:")[1]),
        "velocity": floats(blk[2].split("Velocity:")[1]),
        "omega": floats(blk[3].split("velocity:")[1]),
    }
    cpo = read("tests/dem/combined_periodic_offsets.output")
    three = cpo.split("dim = 3")[1]
    g["combined_periodic_offsets_3d"] = [floats(l.split("::")[1]) for l in three.splitlines() if "(" in l]
    g["contact_on_two_processors_y"] = contact_on_two_processors()
    # two_particles_multiple_contacts_parallel (2 ranks, both spheres on rank 1): force_y on particle 0 every 10th step
    g["two_particles_multiple_contacts"] = [[int(m.group(1)), float(m.group(2))] for m in re.finditer(
        r"at step (\d+) is: (\S+)", read("tests/dem/two_particles_multiple_contacts_parallel.mpirun=2.output"))]
    for k in (1, 2):  # insertion_volume_{1,2}: the 10 inserted positions
        g[f"insertion_volume_{k}"] = [[float(v) for v in l.split("at:")[1].split()]
                                      for l in read(f"tests/dem/insertion_volume_{k}.output").splitlines() if "inserted at" in l]
    # boundary_cells_and_faces: (active cell, deal.II face number) of the 96 boundary faces, in face-number order
    g["boundary_cells_and_faces"] = [[int(m.group(1)), int(m.group(2))] for m in re.finditer(
        r"Cell 2\.(\d+) is on system boundaries \(boundary(\d+)\)", read("tests/dem/boundary_cells_and_faces.output"))]
    g["insertion_plane"] = [[float(v) for v in l.split("at:")[1].split()]
                            for l in read("tests/dem/insertion_plane.output").splitlines() if "inserted at" in l]
    # find_cell_neighbors<3, false> on hyper_cube(-1, 1) refined twice: "2.k" = active cell k of level 2
    fcn = read("tests/dem/find_cell_neighbors.output").split("reciprocal = 1")[0]
    g["find_cell_neighbors"] = [[int(v.split(".")[1]) for v in re.findall(r"2\.\d+", l.split("are:")[1])]
                                for l in fcn.splitlines() if "neighbors of cell" in l]
    g["find_full_cell_neighbors"] = [[int(v) for v in l.split("are:")[1].split()]
                                     for l in read("tests/dem/find_full_cell_neighbors.output").splitlines() if "neighbors of cell" in l]
    g["find_contact_pairs"] = [
        [int(v) for v in re.findall(r"particle (\d+)", l)] for l in read("tests/dem/find_contact_pairs.output").splitlines() if "pair" in l
    ]
    with open(os.path.join(OUT, "unit_goldens.json"), "w") as f:
        json.dump(g, f, indent=0)
    for name in ("packing_in_box.mpirun=1",):
        rows = application_positions(name)
        with open(os.path.join(OUT, name.replace("=", "") + ".json"), "w") as f:
            json.dump({"source": f"applications_tests/lethe-particles/{name}.output", "rows": rows}, f)
        # the matching .prm is an input file of the reference's test-suite (parameters only)
    prm = read("applications_tests/lethe-particles/packing_in_box.prm")
    with open(os.path.join(OUT, "packing_in_box.prm"), "w") as f:
        f.write(prm)
    print("golden fixtures written:", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
