"""Extracts the solid-surface (triangle-mesh wall) known answers of the reference's application
tests into tests/golden/solid_surface_goldens.json (run where /root/reference is mounted).

    python tests/golden/make_solid_goldens.py

Sources: applications_tests/lethe-particles/particle_solid_surface_*.{prm,output} (one sphere
falling on two triangles: face / edge / vertex contacts and the double-contact elimination) and
solid_surfaces_mesh/*.msh. Only numbers are extracted: parameters, mesh coordinates, and the
"Velocity magnitude" column the tests log."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lethe_b200.mesh_io import read_msh_triangles  # noqa: E402

REF = "/root/reference/applications_tests/lethe-particles"
CASES = ["edge_vertex_contact", "CPES_double_edge_contact", "NPES_double_edge_contact", "NPES_double_face_contact"]


def prm_value(txt, key, default=None):
    m = re.search(r"^\s*set\s+" + re.escape(key) + r"\s*=\s*(.*?)\s*$", txt, re.M)
    return m.group(1) if m else default


def main():
    out = {}
    for case in CASES:
        prm = open(f"{REF}/particle_solid_surface_{case}.prm").read()
        log = open(f"{REF}/particle_solid_surface_{case}.output").read()
        mesh = os.path.basename(prm_value(prm, "file name"))
        v, t = read_msh_triangles(f"{REF}/solid_surfaces_mesh/{mesh}")
        lo, hi, _ = prm_value(prm, "grid arguments").split(":")
        out[case] = {
            "dt": float(prm_value(prm, "time step")),
            "t_end": float(prm_value(prm, "time end")),
            "log_frequency": int(prm_value(prm, "log frequency")),
            "g": [float(x) for x in prm_value(prm, "g").split(",")],
            "diameter": float(prm_value(prm, "diameter")),
            "density": float(prm_value(prm, "density particles")),
            "young": float(prm_value(prm, "young modulus particles")),
            "poisson": float(prm_value(prm, "poisson ratio particles")),
            "restitution": float(prm_value(prm, "restitution coefficient particles")),
            "friction": float(prm_value(prm, "friction coefficient particles")),
            "young_wall": float(prm_value(prm, "young modulus wall")),
            "poisson_wall": float(prm_value(prm, "poisson ratio wall")),
            "restitution_wall": float(prm_value(prm, "restitution coefficient wall")),
            "friction_wall": float(prm_value(prm, "friction coefficient wall")),
            "pp_model": prm_value(prm, "particle particle contact force method"),
            "pw_model": prm_value(prm, "particle wall contact force method"),
            "neighborhood_threshold": float(prm_value(prm, "neighborhood threshold")),
            "search_factor": float(prm_value(prm, "dynamic contact search size coefficient")),
            "position": [float(prm_value(prm, "list " + a)) for a in "xyz"],
            "box": [float(lo), float(hi)],
            "refinement": int(prm_value(prm, "initial refinement")),
            "rotation_axis": [float(x) for x in prm_value(prm, "initial rotation axis").split(",")],
            "rotation_angle": float(prm_value(prm, "initial rotation angle", "0")),
            "vertices": v.tolist(),
            "triangles": t.tolist(),
            "velocity_magnitude": [float(m) for m in re.findall(r"Velocity magnitude\s*\|\s*\S+\s*\|\s*(\S+)", log)],
        }
        # the mesh block also has an "initial refinement" (0): the first one in the file is the solid's
        refs = re.findall(r"set initial refinement\s*=\s*(\d+)", prm)
        out[case]["refinement"] = max(int(r) for r in refs)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "solid_surface_goldens.json"), "w") as f:
        json.dump(out, f, indent=1)
    for k, v in out.items():
        print(k, len(v["velocity_magnitude"]), "samples", v["box"], v["refinement"], v["rotation_angle"])


if __name__ == "__main__":
    main()
