"""Manufactured solutions used by the reference's own tests (inputs of the golden files)."""
import numpy as np

pi = np.pi
s, c = np.sin, np.cos


def forcing_2d(x):
    """MMSSineForcingFunction, tests/solvers/restart_01.cc:42-63 (exact pressure x^2+y^2)."""
    a, X, Y = pi, x[:, 0], x[:, 1]
    f0 = (2 * a * a * (-s(a * X) * s(a * X) + c(a * X) * c(a * X)) * s(a * Y) * c(a * Y)
          - 4 * a * a * s(a * X) * s(a * X) * s(a * Y) * c(a * Y) - 2.0 * X) * (-1.) \
        + a * s(a * X) ** 3 * s(a * Y) ** 2 * c(a * X)
    f1 = (2 * a * a * (s(a * Y) * s(a * Y) - c(a * Y) * c(a * Y)) * s(a * X) * c(a * X)
          + 4 * a * a * s(a * X) * s(a * Y) * s(a * Y) * c(a * X) - 2.0 * Y) * (-1) \
        + a * s(a * X) ** 2 * s(a * Y) ** 3 * c(a * Y)
    return np.stack([f0, f1], axis=1)


def exact_2d(x):
    """ExactSolutionMMS, tests/solvers/restart_01.cc:19-29 (pressure component left 0)."""
    a, X, Y = pi, x[:, 0], x[:, 1]
    return np.stack([s(a * X) ** 2 * c(a * Y) * s(a * Y), -c(a * X) * s(a * X) * s(a * Y) ** 2,
                     0 * X], axis=1)


def forcing_3d(X):
    """applications_tests/gls_navier_stokes_3d/mms3d_gls.prm, subsection source term."""
    x, y, z = X[:, 0], X[:, 1], X[:, 2]
    f0 = 2 * pi * pi * (-3 * c(2 * pi * x) + 2.) * s(pi * y) * s(pi * z) * c(pi * y) * c(pi * z) \
        + pi * (2 * (c(pi * y) ** 2) - (c(pi * z) ** 2)) * (s(pi * x) ** 3) * (s(pi * y) ** 2) \
        * (s(pi * z) ** 2) * c(pi * x)
    f1 = 2 * pi * pi * (-3 * c(2 * pi * y) + 2) * s(pi * x) * s(pi * z) * c(pi * x) * c(pi * z) \
        + pi * (2 * (c(pi * x) ** 2) - (c(pi * z) ** 2)) * (s(pi * x) ** 2) * (s(pi * y) ** 3) \
        * (s(pi * z) ** 2) * c(pi * y)
    f2 = 4 * pi * pi * (3 * c(2 * pi * z) - 2) * s(pi * x) * s(pi * y) * c(pi * x) * c(pi * y) \
        + 2 * pi * ((c(pi * x) ** 2) + (c(pi * y) ** 2)) * (s(pi * x) ** 2) * (s(pi * y) ** 2) \
        * (s(pi * z) ** 3) * c(pi * z)
    return np.stack([f0, f1, f2], axis=1)


def exact_3d(X):
    """applications_tests/gls_navier_stokes_3d/mms3d_gls.prm, subsection analytical solution."""
    x, y, z = X[:, 0], X[:, 1], X[:, 2]
    return np.stack([s(pi * x) ** 2 * c(pi * y) * s(pi * y) * c(pi * z) * s(pi * z),
                     c(pi * x) * s(pi * x) * s(pi * y) ** 2 * c(pi * z) * s(pi * z),
                     -2 * c(pi * x) * s(pi * x) * c(pi * y) * s(pi * y) * s(pi * z) ** 2,
                     0 * x], axis=1)


def forcing_mms2d(x):
    """applications_tests/gls_navier_stokes_2d/mms2d_gls.prm, subsection source term
    (exact pressure sin(pi x) + sin(pi y))."""
    X, Y = x[:, 0], x[:, 1]
    f0 = (2 * pi * pi * (-s(pi * X) * s(pi * X) + c(pi * X) * c(pi * X)) * s(pi * Y) * c(pi * Y)
          - 4 * pi * pi * s(pi * X) * s(pi * X) * s(pi * Y) * c(pi * Y) - pi * c(pi * X)) * (-1.) \
        + pi * s(pi * X) ** 3 * s(pi * Y) ** 2 * c(pi * X)
    f1 = (2 * pi * pi * (s(pi * Y) * s(pi * Y) - c(pi * Y) * c(pi * Y)) * s(pi * X) * c(pi * X)
          + 4 * pi * pi * s(pi * X) * s(pi * Y) * s(pi * Y) * c(pi * X) - pi * c(pi * Y)) * (-1) \
        + pi * s(pi * X) ** 2 * s(pi * Y) ** 3 * c(pi * Y)
    return np.stack([f0, f1], axis=1)


def exact_mms2d(x):
    """applications_tests/gls_navier_stokes_2d/mms2d_gls.prm, subsection analytical solution."""
    X, Y = x[:, 0], x[:, 1]
    return np.stack([s(pi * X) ** 2 * c(pi * Y) * s(pi * Y), -c(pi * X) * s(pi * X) * s(pi * Y) ** 2,
                     s(pi * X) + s(pi * Y)], axis=1)
