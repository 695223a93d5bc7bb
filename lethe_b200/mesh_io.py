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
