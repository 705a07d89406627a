"""CPU-only checks of libamgb.so: the library loads, exports every symbol
include/amgb.h declares, its host-side setup (generators, interpolation maps,
level sizes) is bit-identical to the oracle, and compute entry points fail
loudly without a device (no CPU fallback)."""
import importlib
import os
import re

import numpy as np
import pytest

import oracle as O

amg = importlib.import_module("algebraic-multigrid_b200")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported_and_bound():
    header = open(os.path.join(ROOT, "include", "amgb.h")).read()
    declared = set(re.findall(r"\b(amgb_[a-z_A-Z0-9]+)\s*\(", header))
    declared -= {"amgb_matrix", "amgb_hierarchy", "amgb_options"}
    L = amg.lib()
    for name in sorted(declared):
        assert hasattr(L, name), "libamgb.so does not export %s" % name
    assert declared == set(amg.SIGNATURES), declared ^ set(amg.SIGNATURES)
    assert L.amgb_version() >= 100


@pytest.mark.parametrize("n,eps", [(2, 1.0), (35, 1.0), (64, 1.0), (33, 1e-3)])
def test_generators_bit_identical_to_oracle(n, eps):
    A = amg.Grid.laplacian(n, eps)
    Ao = O.laplacian(n, eps)
    colptr, rowidx, val = Ao.arrays()
    np.testing.assert_array_equal(A.colptr, colptr)
    np.testing.assert_array_equal(A.rowidx, rowidx)
    assert A.val.tobytes() == val.tobytes()
    assert amg.Grid.rhs(n).tobytes() == O.rhs(n).tobytes()


def test_grid_helpers():
    assert amg.Grid.grid_spacing_h(2) == 2.0 / 3.0
    assert amg.Grid.points_n_from_grid_spacing_h(amg.Grid.grid_spacing_h(2)) == 2  # testlib.cpp:60-62


@pytest.mark.parametrize("nh", [7, 24, 1225, 612, 8, 3, 2])
def test_interpolation_maps_bit_exact(nh):
    nH = amg.lib().amgb_n_H_dofs_from_n_h_dofs(nh)
    assert nH == O.n_H_from_n_h(nh)
    if nH < 1:
        return
    li = amg.LinearInterpolator(2)
    li.make_operators(nh, nH, 0)
    P, R = li.get_P(0), li.get_R(0)
    Po = O.make_P(nh, nH)
    Ro = Po.transpose()
    for mine, ref in ((P, Po), (R, Ro)):
        c, r, v = ref.arrays()
        np.testing.assert_array_equal(mine.colptr, c)
        np.testing.assert_array_equal(mine.rowidx, r)
        assert mine.val.tobytes() == v.tobytes()


def test_level_size_rule():
    sizes = [4097 * 4097]
    for _ in range(15):
        sizes.append(amg.lib().amgb_n_H_dofs_from_n_h_dofs(sizes[-1]))
    assert sizes == O.level_sizes(4097 * 4097, 16)


@pytest.mark.skipif(amg.device_count() > 0, reason="checks the no-device behaviour")
def test_no_cpu_fallback():
    A = amg.Grid.laplacian(4)
    b = amg.Grid.rhs(4)
    with pytest.raises(amg.AmgbError) as e:
        amg.DeviceMatrix(A)
    assert e.value.code == amg.ECUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(amg.AmgbError) as e:
        amg.Multigrid(None, amg.SparseGaussSeidel(), A, b, 2, 1e-9, 1, 1)
    assert e.value.code == amg.ECUDA


def test_ctor_validation_precedes_device_use():
    # multigrid.hpp:165-178 / testlib.cpp:131-144: both conditions, in this order
    A = amg.Grid.laplacian(2)
    b = amg.Grid.rhs(2)
    with pytest.raises(amg.InvalidArgument, match="compute_error_every_n_iters"):
        amg.Multigrid(None, amg.SparseGaussSeidel(), A, b, 8, 1e-9, 100, 10)
    with pytest.raises(amg.InvalidArgument, match="same number of degrees of freedom"):
        amg.Multigrid(None, amg.SparseGaussSeidel(), A, np.zeros(5), 8, 1e-9, 5, 10)
    with pytest.raises(amg.InvalidArgument, match="compute_error_every_n_iters"):
        amg.Multigrid(None, amg.SparseGaussSeidel(), A, np.zeros(5), 8, 1e-9, 100, 10)
