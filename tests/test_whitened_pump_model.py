"""CPU: the identity the dense-metric pump (turing.jl_b200/csrc/dense_pump.cuh) rests on, checked in
numpy on the oracle's log density: with M^-1 = U'U (U upper triangular), AdvancedHMC's dense-metric
leapfrog in (theta, r) -- drift theta += eps M^-1 r, kick r += eps/2 grad -- is the unit-metric leapfrog
in phi = U^-T theta, p = U r with the gradient U grad_theta, and the quantities the NUTS tree consumes
agree: kinetic energy r'M^-1 r / 2 = p'p / 2, the generalised U-turn products rho . M^-1 r = rho_p . p,
and the momentum draw r = U \\ z is p = z."""
import numpy as np

import oracle as O
import turing_jl_b200 as T


def test_dense_metric_dynamics_is_unit_metric_dynamics_in_whitened_coordinates(built):
    model = T.models.synthetic_logistic(200, 9, seed=3)
    om = O.OracleModel(O.LOGISTIC, D=model.D, N=model.N, X=model.X, y=model.y, hyper=model.hyper)
    D = model.D
    rng = np.random.default_rng(0)
    A = rng.standard_normal((D, D))
    Minv = A @ A.T / D + 0.1 * np.eye(D)
    U = np.linalg.cholesky(Minv).T                     # upper, M^-1 = U'U
    np.testing.assert_allclose(U.T @ U, Minv, rtol=1e-13)
    theta = rng.standard_normal(D) * 0.3
    z = rng.standard_normal(D)
    r = np.linalg.solve(U, z)                          # AdvancedHMC: r = U \ z, cov(r) = M
    phi, p = np.linalg.solve(U.T, theta), U @ r
    np.testing.assert_allclose(p, z, rtol=1e-12)       # the momentum of the whitened chain is z itself
    eps, n = 0.07, 25
    _, g = om.lpgrad(theta)
    gp = U @ g
    rho, rho_p = r.copy(), p.copy()
    for _ in range(n):
        # dense metric, theta space
        r = r + 0.5 * eps * g
        theta = theta + eps * (Minv @ r)
        lp, g = om.lpgrad(theta)
        r = r + 0.5 * eps * g
        rho += r
        # unit metric, phi space; the family's gradient is evaluated at theta = U'phi
        p = p + 0.5 * eps * gp
        phi = phi + eps * p
        lp2, g2 = om.lpgrad(U.T @ phi)
        gp = U @ g2
        p = p + 0.5 * eps * gp
        rho_p += p
        np.testing.assert_allclose(U.T @ phi, theta, rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(lp2, lp, rtol=1e-12)
        np.testing.assert_allclose(0.5 * p @ p, 0.5 * r @ (Minv @ r), rtol=1e-11)       # kinetic energy
        np.testing.assert_allclose(rho_p @ p, rho @ (Minv @ r), rtol=1e-10, atol=1e-10)  # U-turn product
    # a new metric re-expresses the same state: phi' = U'^-T theta, grad_phi' = U' grad_theta
    B = rng.standard_normal((D, D))
    U2 = np.linalg.cholesky(B @ B.T / D + 0.05 * np.eye(D)).T
    g_theta = np.linalg.solve(U, gp)                   # back substitution, as k_dense_post does
    np.testing.assert_allclose(g_theta, g, rtol=1e-10)
    np.testing.assert_allclose(U2.T @ np.linalg.solve(U2.T, theta), theta, rtol=1e-12)
