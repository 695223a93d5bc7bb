"""Reader for the triangle surfaces of `subsection solid objects` (gmsh .msh, format 4.1 and
2.2, ASCII) — what GridIn::read_msh gives SerialSolid<2,3>::setup_triangulation
(source/core/serial_solid.cc:163-175): vertices in node order, triangles in element order."""
from __future__ import annotations

import numpy as np


def read_msh_triangles(path):
    with open(path) as f:
        lines = [ln.strip() for ln in f]
    sections = {}
    k = 0
    while k < len(lines):
        if lines[k].startswith("$") and not lines[k].startswith("$End"):
            name = lines[k][1:]
            end = lines.index("$End" + name, k)
            sections[name] = lines[k + 1:end]
            k = end
        k += 1
    version = float(sections["MeshFormat"][0].split()[0])
    nodes, tris = {}, []
    if version >= 4.0:
        body = sections["Nodes"]
        n_blocks = int(body[0].split()[0])
        k = 1
        for _ in range(n_blocks):
            n_in_block = int(body[k].split()[3])
            tags = [int(body[k + 1 + i]) for i in range(n_in_block)]
            for i, tag in enumerate(tags):
                nodes[tag] = [float(v) for v in body[k + 1 + n_in_block + i].split()[:3]]
            k += 1 + 2 * n_in_block
        body = sections["Elements"]
        n_blocks = int(body[0].split()[0])
        k = 1
        for _ in range(n_blocks):
            _, _, etype, n_in_block = (int(v) for v in body[k].split())
            for i in range(n_in_block):
                parts = [int(v) for v in body[k + 1 + i].split()]
                if etype == 2:
                    tris.append(parts[1:4])
            k += 1 + n_in_block
    else:
        body = sections["Nodes"]
        for ln in body[1:1 + int(body[0])]:
            parts = ln.split()
            nodes[int(parts[0])] = [float(v) for v in parts[1:4]]
        body = sections["Elements"]
        for ln in body[1:1 + int(body[0])]:
            parts = [int(v) for v in ln.split()]
            if parts[1] == 2:
                tris.append(parts[3 + parts[2]:3 + parts[2] + 3])
    tags = sorted(nodes)
    index = {tag: i for i, tag in enumerate(tags)}
    vertices = np.array([nodes[t] for t in tags], dtype=np.float64)
    triangles = np.array([[index[v] for v in t] for t in tris], dtype=np.uint32)
    return vertices, triangles


def dealii_simplex_surface(grid_type, grid_arguments, initial_refinement=0):
    """`mesh type = dealii` with `simplex = true` for a solid surface (serial_solid.cc:176-196):
    GridGenerator::generate_from_name_and_arguments on a Triangulation<2,3>, `initial refinement`
    global refinements, flatten, GridGenerator::convert_hypercube_to_simplex_mesh. deal.II is an
    external dependency of the reference (not vendored): restated here are hyper_cube /
    hyper_rectangle in the z = 0 plane and deal.II's published 2-D conversion, every quadrilateral
    -> 8 triangles over its 4 corners (0-3, lexicographic), 4 edge midpoints (4: x-low, 5: x-high,
    6: y-low, 7: y-high) and its centre (8). Pinned by the reference's `solid_surface.output`."""
    args = [a.strip() for a in grid_arguments.split(":")]
    if grid_type == "hyper_cube":
        lo, hi = float(args[0]), float(args[1])
        p1, p2 = (lo, lo), (hi, hi)
    elif grid_type == "hyper_rectangle":
        p1 = tuple(float(v) for v in args[0].split(","))[:2]
        p2 = tuple(float(v) for v in args[1].split(","))[:2]
    else:
        raise ValueError(f"solid surfaces: dealii grid type `{grid_type}` is not generated here (hyper_cube, hyper_rectangle)")
    n = 1 << int(initial_refinement)
    hx, hy = (p2[0] - p1[0]) / n, (p2[1] - p1[1]) / n
    table = ((0, 6, 4), (8, 4, 6), (8, 6, 5), (1, 5, 6), (2, 4, 7), (8, 7, 4), (8, 5, 7), (3, 7, 5))
    index, vertices, triangles = {}, [], []

    def vertex(i2, j2):  # half-cell lattice coordinates
        key = (i2, j2)
        if key not in index:
            index[key] = len(vertices)
            vertices.append((p1[0] + 0.5 * i2 * hx, p1[1] + 0.5 * j2 * hy, 0.0))
        return index[key]

    def morton(k):
        i = j = 0
        for b in range(int(initial_refinement)):
            i |= ((k >> (2 * b)) & 1) << b
            j |= ((k >> (2 * b + 1)) & 1) << b
        return i, j

    cells = [morton(k) for k in range(n * n)]
    for i, j in cells:  # the quadrilateral mesh's own vertices come first
        for dj in (0, 2):
            for di in (0, 2):
                vertex(2 * i + di, 2 * j + dj)
    for i, j in cells:
        local = [vertex(2 * i, 2 * j), vertex(2 * i + 2, 2 * j), vertex(2 * i, 2 * j + 2), vertex(2 * i + 2, 2 * j + 2),
                 vertex(2 * i, 2 * j + 1), vertex(2 * i + 2, 2 * j + 1), vertex(2 * i + 1, 2 * j), vertex(2 * i + 1, 2 * j + 2),
                 vertex(2 * i + 1, 2 * j + 1)]
        triangles += [[local[a], local[b], local[c]] for a, b, c in table]
    return np.array(vertices, dtype=np.float64), np.array(triangles, dtype=np.uint32)
