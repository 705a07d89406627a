"""Chunk geometry and edge classification of the streaming legs (stream_leg_api.hpp), on the CPU.
The fused halo push is only correct if (1) the chunks cover every line exactly once, (2) every chunk
whose stages can read a ghost row is an edge chunk on that side (it waits for the neighbour's flag), and
(3) every chunk that stores a row the neighbour receives is an edge chunk on that side (it pushes and
is counted before the flag is raised).  Checked by brute force over many shapes."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "sleg_geometry_host.cpp")
HDR = os.path.join(ROOT, "algebraic-multigrid_b200", "csrc", "stream_leg_api.hpp")
LIB = os.path.join(ROOT, "tests", "cpp", "libsleg_geometry_host.so")
CUDA_INC = "/usr/local/cuda/include"


def lib():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-Wall", "-I", CUDA_INC, "-x", "c++", SRC,
                               "-o", LIB])
    L = C.CDLL(LIB)
    L.sleg_geometry.argtypes = [C.c_int] * 8 + [np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")]
    L.sleg_chunk_begin.restype = C.c_int
    L.sleg_chunk_lines.restype = C.c_int
    return L


@pytest.mark.parametrize("edge_half", [0, 1])
def test_chunks_cover_every_line_once(edge_half):
    L = lib()
    rng = np.random.default_rng(0)
    for _ in range(300):
        n_lines = int(rng.integers(1, 5000))
        n_strips = int(rng.integers(1, 200))
        chunks = int(rng.integers(1, 400))
        out = np.zeros(8, np.int32)
        L.sleg_geometry(n_lines, 64, n_strips, chunks, edge_half, 2, 0, n_lines * 64, out)
        LJ, LJe, nc, nw = out[:4]
        assert nw == nc * n_strips and 1 <= nc <= max(chunks, 2) + 1
        covered = 0
        for c in range(nc):
            b, ln = L.sleg_chunk_begin(n_lines, n_strips, chunks, edge_half, c), L.sleg_chunk_lines(n_lines, n_strips, chunks, edge_half, c)
            assert b == covered, (n_lines, chunks, c, b, covered)
            covered += ln
        assert covered >= n_lines and covered - LJ < n_lines      # the last chunk may run past the end, never a whole chunk


@pytest.mark.parametrize("edge_half", [0, 1])
def test_edge_chunks_contain_every_ghost_reader_and_every_pusher(edge_half):
    L = lib()
    rng = np.random.default_rng(1)
    for _ in range(400):
        m = int(rng.integers(3, 300))
        halo = 3 * (m + 1) + 4                      # ghost rows on each side of a sharded window
        n_own = int(rng.integers(4 * halo, 40 * halo))
        own_begin, own_end = halo, halo + n_own
        n = n_own + 2 * halo
        n_lines = (n + m - 1) // m
        NS = int(rng.integers(1, 4))
        PF = 2
        n_strips = (m + 23) // 24
        chunks = int(rng.integers(1, 200))
        push_lo_end = own_begin + int(rng.integers(0, min(halo, n_own) + 1))     # rows [own_begin, push_lo_end) go down
        push_hi_begin = own_end - int(rng.integers(0, min(halo, n_own) + 1))     # rows [push_hi_begin, own_end) go up
        out = np.zeros(8, np.int32)
        L.sleg_geometry(n_lines, m, n_strips, chunks, edge_half, NS, max(own_begin, push_lo_end), min(own_end, push_hi_begin), out)
        nc, e_lo, e_hi0 = int(out[2]), int(out[4]), int(out[5])
        assert out[6] == e_lo * n_strips and out[7] == (nc - e_hi0) * n_strips
        for c in range(nc):
            j0 = L.sleg_chunk_begin(n_lines, n_strips, chunks, edge_half, c)
            j1 = min(j0 + L.sleg_chunk_lines(n_lines, n_strips, chunks, edge_half, c), n_lines)
            # rows whose values can reach a stored result of this chunk: stage inputs of lines j0 - NS .. j1 - 1 + NS,
            # lanes up to 4 rows beyond the line on either side
            first_used = (j0 - NS) * m - 4
            last_used = (j1 - 1 + NS) * m + m + 4
            stores_lo, stores_hi = max(j0 * m, own_begin), min(j1 * m, own_end)
            reads_lower_ghost = first_used < own_begin and stores_hi > stores_lo
            reads_upper_ghost = last_used >= own_end and stores_hi > stores_lo
            pushes_down = stores_lo < push_lo_end and stores_hi > stores_lo
            pushes_up = stores_hi > push_hi_begin and stores_hi > stores_lo
            if reads_lower_ghost or pushes_down:
                assert c < e_lo, (m, n_own, NS, chunks, c, e_lo)
            if reads_upper_ghost or pushes_up:
                assert c >= e_hi0, (m, n_own, NS, chunks, c, e_hi0)
        del PF
