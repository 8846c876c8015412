"""``.mcl`` snapshot files — the reference's fixture format for one scan + particle set
(include/tsdf_localization/util/mcl_file.h:21-32, src/util/mcl_file.cpp:14-113; written by snap_shot_node, replayed by
num_particles_eval / cuda_test_eval / snap_vis_node). Whitespace-separated text:

    P                       number of points
    P x "x y z"             points (sensor frame)
    P x "ring"              ring of every point
    N                       number of particles
    N x "x y z roll pitch yaw  weight"
    16 floats               scanner->robot transform, row-major
    "x y z q1 q2 q3 q4"     reference pose

Numbers are formatted like the reference's ``ostream << float`` (``%g``, 6 significant digits) so files written here are
byte-identical to the reference's for the same values; reading accepts anything ``istream >> float`` accepts.
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path

import numpy as np


@dataclass
class MCLSnapshot:
    points: np.ndarray        # float32 [P, 3]
    rings: np.ndarray         # int32 [P]
    particles: np.ndarray     # float32 [N, 7]  x y z roll pitch yaw weight
    tf: np.ndarray            # float32 [16]
    pose: np.ndarray          # float32 [7]  x y z q1 q2 q3 q4


def _g(v) -> str:
    return "%g" % float(v)


class MCLFile:
    def __init__(self, file_name):
        self.name_ = str(file_name)

    def write(self, points, rings, particles, tf, x, y, z, q_1, q_2, q_3, q_4) -> None:
        pts = np.asarray(points, dtype=np.float32).reshape(-1, 3)
        rg = np.asarray(rings).reshape(-1)
        ps = np.asarray(particles, dtype=np.float32).reshape(-1, 7)
        tfm = np.asarray(tf, dtype=np.float32).reshape(-1)
        if len(rg) != len(pts) or len(tfm) != 16:
            raise ValueError("rings must match points; tf must hold 16 values")
        out = [f"{len(pts)}\n\n"]
        out += [f"{_g(p[0])} {_g(p[1])} {_g(p[2])}\n" for p in pts]
        out += [f"{int(r)}\n" for r in rg]
        out.append(f"{len(ps)}\n")
        out += [" ".join(_g(v) for v in p[:6]) + "  " + _g(p[6]) + "\n" for p in ps]
        out.append("\n")
        out.append("".join(_g(v) + " " for v in tfm))
        out.append("\n")
        out.append(" ".join(_g(np.float32(v)) for v in (x, y, z, q_1, q_2, q_3, q_4)))
        try:
            Path(self.name_).write_text("".join(out))
        except OSError as e:
            raise OSError("Error while opening file for writing") from e

    def read(self) -> MCLSnapshot:
        try:
            tokens = Path(self.name_).read_text().split()
        except OSError as e:
            raise OSError("Error while opening file for reading") from e
        pos = 0

        def take(n, dtype):
            nonlocal pos
            if pos + n > len(tokens):
                raise ValueError("Error: Could not read mcl data from file")
            try:
                vals = np.array(tokens[pos:pos + n], dtype=np.float64).astype(dtype) if dtype != np.int32 else \
                    np.array([int(t) for t in tokens[pos:pos + n]], dtype=np.int32)
            except ValueError as e:
                raise ValueError("Error: Could not read mcl data from file") from e
            pos += n
            return vals

        try:
            n_points = int(tokens[0])
        except (IndexError, ValueError) as e:
            raise ValueError("Error: Could not read mcl data from file") from e
        pos = 1
        points = take(3 * n_points, np.float32).reshape(n_points, 3)
        rings = take(n_points, np.int32)
        n_particles = int(take(1, np.int32)[0])
        particles = take(7 * n_particles, np.float32).reshape(n_particles, 7)
        tf = take(16, np.float32)
        pose = take(7, np.float32)
        return MCLSnapshot(points, rings, particles, tf, pose)
