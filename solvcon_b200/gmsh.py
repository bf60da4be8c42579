"""Gmsh MSH 2.2 (ASCII) reader -> ``StaticMesh`` -- the reference's ``solvcon.Gmsh`` (``cpp/solvcon/inout/gmsh.hpp``,
``gmsh.cpp``; binding ``Gmsh(data).to_block()``), the only on-disk mesh format its multidim tests use
(``tests/test_gmsh.py``, ``tests/test_euler_multidim.py:191-224``).  SURVEY.md 8(f) row f3.

Same behaviour, section by section (``gmsh.cpp:10-64``, ``gmsh.hpp:115-262``):

* a finite-state machine walks ``$MeshFormat`` -> [``$PhysicalNames``] -> ``$Nodes`` -> ``$Elements``; a ``$`` line
  that is not one of those keywords, or that arrives out of order, silently ENDS the parse; if one of the three
  mandatory sections was never entered the constructor raises ``ValueError("Incorrect msh file format.")``;
* only format version 2.2 is accepted (``ValueError``);
* node and element numbers are 1-based in the file and 0-based in the mesh; of a high-order element only the corner
  nodes are kept, in solvcon's local order (``GmshElementDef::by_id``: the prism swaps nodes 1<->2 and 4<->5);
* ``to_block()`` keeps only the elements of the highest dimension as cells (points and boundary lines merely tag
  physical groups), then runs ``build_interior(True)``, ``build_boundary()``, ``build_ghost()`` -- here through the
  native table builder behind :class:`solvcon_b200.mesh.StaticMesh`, so the result can go straight to the device
  ``EulerCore``.
"""
from __future__ import annotations

import numpy as np

from .mesh import make_mesh

__all__ = ["Gmsh"]

# gmsh element type -> (dimension, solvcon cell type, node order)   gmsh.hpp:288-333
_PRISM = (0, 2, 1, 3, 5, 4)
_ELEMENT_DEFS = {
    1: (1, 2, (0, 1)), 2: (2, 4, (0, 1, 2)), 3: (2, 3, (0, 1, 2, 3)), 4: (3, 6, (0, 1, 2, 3)),
    5: (3, 5, tuple(range(8))), 6: (3, 7, _PRISM), 7: (3, 8, tuple(range(5))), 8: (1, 2, (0, 1)),
    9: (2, 4, (0, 1, 2)), 10: (2, 3, (0, 1, 2, 3)), 11: (3, 6, (0, 1, 2, 3)), 12: (3, 5, tuple(range(8))),
    13: (3, 7, _PRISM), 14: (3, 8, tuple(range(5))), 15: (0, 1, (0,)), 16: (2, 3, (0, 1, 2, 3)),
    17: (3, 5, tuple(range(8))), 18: (3, 7, _PRISM), 19: (3, 8, tuple(range(5))), 20: (2, 4, (0, 1, 2)),
    21: (2, 4, (0, 1, 2)), 22: (2, 4, (0, 1, 2)), 23: (2, 4, (0, 1, 2)), 24: (2, 4, (0, 1, 2)), 25: (2, 4, (0, 1, 2)),
    26: (1, 2, (0, 1)), 27: (1, 2, (0, 1)), 28: (1, 2, (0, 1)), 29: (3, 6, (0, 1, 2, 3)), 30: (3, 6, (0, 1, 2, 3)),
    31: (3, 6, (0, 1, 2, 3)), 92: (3, 5, tuple(range(8))), 93: (3, 5, tuple(range(8))),
}
_CLMND = 8


class Gmsh:
    def __init__(self, data):
        if isinstance(data, (bytes, bytearray)):
            data = bytes(data).decode("latin-1")
        self._lines = data.replace("\r", "").split("\n")
        if self._lines and self._lines[-1] == "":
            self._lines.pop()   # std::getline yields no empty record after the final newline
        self._pos = 0
        self._state = "BEGIN"
        self.msh_ver = 0.0
        self.msh_file_type = 0
        self.msh_data_size = 0
        self.nodes = np.zeros((0, 3))
        self.eltype, self.elgrp, self.elgeo, self.eldim, self.elnds = [], [], [], [], []
        entered = {"meta": False, "node": False, "element": False}
        handlers = {"$MeshFormat": (self._load_meta, "meta"), "$Nodes": (self._load_nodes, "node"),
                    "$Elements": (self._load_elements, "element"), "$PhysicalNames": (self._load_physical, None)}
        while self._pos < len(self._lines):
            line = self._next()
            if "$" in line:
                if line not in handlers or not self._valid_transition(line):
                    break
                fn, key = handlers[line]
                fn()
                if key:
                    entered[key] = True
        if not all(entered.values()):
            raise ValueError("Incorrect msh file format.")

    # -- line cursor ------------------------------------------------------------------------------------
    def _next(self):
        line = self._lines[self._pos]
        self._pos += 1
        return line

    def _section_lines(self):
        """Lines up to (and consuming) the next one that contains '$'; returns (lines, terminator or None)."""
        out = []
        while self._pos < len(self._lines):
            line = self._next()
            if "$" in line:
                return out, line
            out.append(line)
        return out, None

    def _valid_transition(self, s):      # gmsh.hpp:115-131
        if self._state == "BEGIN":
            return s == "$MeshFormat"
        if self._state in ("META_END", "PHYSICAL_NAME_END"):
            return s in ("$PhysicalNames", "$Nodes")
        if self._state == "NODE_END":
            return s == "$Elements"
        return False

    # -- sections ---------------------------------------------------------------------------------------
    def _load_meta(self):                # gmsh.hpp:133-157
        lines, term = self._section_lines()
        for line in lines:
            tok = [t for t in line.split(" ") if t]
            self.msh_ver = float(tok[0])
            if self.msh_ver != 2.2:
                raise ValueError(f"solvcon does not support msh file ver {self.msh_ver:g}.")
            self.msh_file_type = int(tok[1])
            self.msh_data_size = int(tok[2])
        if term == "$EndMeshFormat":
            self._state = "META_END"

    def _load_physical(self):            # gmsh.hpp:161-171 (names are skipped by the reference as well)
        _, term = self._section_lines()
        if term == "$EndPhysicalNames":
            self._state = "PHYSICAL_NAME_END"

    def _load_nodes(self):               # gmsh.hpp:173-195
        nnode = int(self._next())
        lines, term = self._section_lines()
        nds = np.zeros((nnode, 3))
        if lines:
            tab = np.array([[float(t) for t in line.split(" ") if t][:4] for line in lines])
            nds[tab[:, 0].astype(np.int64) - 1] = tab[:, 1:4]
        self.nodes = nds
        if term == "$EndNodes":
            self._state = "NODE_END"

    def _load_elements(self):            # gmsh.hpp:197-262
        nelement = int(self._next())
        lines, term = self._section_lines()
        for line in lines[:nelement]:
            tok = [int(t) for t in line.split(" ") if t]
            dim, cltpn, order = _ELEMENT_DEFS[tok[1]]
            ntag = tok[2]
            tags = tok[3:3 + ntag]
            nds = tok[3 + ntag:]
            loc = [0] * len(order)
            for i, slot in enumerate(order):
                loc[slot] = nds[i] - 1
            self.eltype.append(cltpn)
            self.elgrp.append(tags[0])
            self.elgeo.append(tags[1])
            self.eldim.append(dim)
            self.elnds.append(loc)
        if term == "$EndElements":
            self._state = "ELEMENT_END"

    # -- mesh -------------------------------------------------------------------------------------------
    def to_block(self):
        """``Gmsh::to_block`` (gmsh.cpp:66-90): cells = the elements of the highest dimension."""
        topdim = max(self.eldim)
        cells = [i for i, d in enumerate(self.eldim) if d == topdim]
        cltpn = np.array([self.eltype[i] for i in cells], dtype=np.int32)
        clnds = np.full((len(cells), _CLMND + 1), -1, dtype=np.int32)
        for icl, iel in enumerate(cells):
            nds = self.elnds[iel]
            clnds[icl, 0] = len(nds)
            clnds[icl, 1:1 + len(nds)] = nds
        # unused clnds slots are left as allocated by the reference (uninitialised); here they are -1
        return make_mesh(topdim, np.ascontiguousarray(self.nodes[:, :topdim]), cltpn, clnds, build_edge=True)
