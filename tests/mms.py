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


# ---- Taylor-Couette (applications_tests/gls_navier_stokes_2d/taylorcouette_gls.prm) ----
def shell_mapping(p):
    """Parameter box (r, theta) -> the annulus: GridGenerator::hyper_shell with its
    SphericalManifold, uniformly refined (new vertices and MappingQ support points at the polar
    midpoints)."""
    return np.stack([p[:, 0] * np.cos(p[:, 1]), p[:, 0] * np.sin(p[:, 1])], axis=1)


def couette_inner_wall(x):
    """bc 0 of taylorcouette_gls.prm:72-83: u = -y, v = x (inner cylinder at unit angular speed)."""
    return np.stack([-x[:, 1], x[:, 0]], axis=1)


def couette_exact(x):
    """taylorcouette_gls.prm:36-43 (eta = ri = 0.25)."""
    eta, ri = 0.25, 0.25
    A, B = -(eta * eta) / (1 - eta * eta), ri * ri / (1 - eta * eta)
    r, th = np.hypot(x[:, 0], x[:, 1]), np.arctan2(x[:, 1], x[:, 0])
    ut = A * r + B / r
    return np.stack([-np.sin(th) * ut, np.cos(th) * ut,
                     A * A * r * r / 2 + 2 * A * B * np.log(r) - 0.5 * B * B / (r * r)], axis=1)


def couette_mesh(oracle, refinement):
    """hyper_shell(0.25, 1, 4 cells) refined `refinement` times: 4 * 2^r cells around, 2^r across."""
    k = 2 ** refinement
    return oracle.BoxMesh(2, (k, 4 * k), 2, 2, lo=(0.25, 0.0), hi=(1.0, 2 * np.pi),
                          bcs={1: ("noslip",), 0: ("function", couette_inner_wall)}, periodic=(1,),
                          mapping=shell_mapping)
