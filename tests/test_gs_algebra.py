"""CPU checks of the algebra the Gauss-Seidel pipeline relies on (mpmcxx_b200/csrc/kernels_gs.cuh), in numpy:
the block walk as dmu = L^-1 (alpha E_s - mu_old - alpha acc) equals the reference's site-by-site update
(contract_dipoles with polar_gs, src/System.Energy.cpp:3570-3595) with the running contraction the engine keeps, ef_induced is
recovered from mu, and the packed layouts of k_gs_inverse / its in-place scratch are consistent."""
import numpy as np

B = 64           # kGsB
N = 3 * B        # kGsN
KINV = 9 * (B * (B - 1) // 2)


def inv_rows(j):
    return N - 3 * (j + 1)


def inv_off(j):
    return 3 * (N - 3) * j - 9 * (j * (j - 1) // 2)


def row_off(r):
    i = r // 3
    return 9 * (i * (i - 1) // 2) + (r - 3 * i) * 3 * i


def t_off(i):
    return KINV - 3 * (B * (B - 1) - i * (i - 1))


def test_packed_layouts():
    # the solver's layout: 21 tiles of 32 x 32 covering the lower triangle of the 192 x 192 inverse, tile (I, J) at I (I + 1) / 2 + J
    tiles = sorted(I * (I + 1) // 2 + J for I in range(N // 32) for J in range(I + 1))
    assert tiles == list(range(21)) and 21 * 32 * 32 * 8 == 172032 and (21 * 32 * 32 * 8 // 3) % 16 == 0
    # every strictly-lower-by-site entry (r, c), c < 3 (r // 3), falls in exactly one stored tile
    covered = np.zeros((N, N), bool)
    for I in range(N // 32):
        for J in range(I + 1):
            covered[32 * I:32 * I + 32, 32 * J:32 * J + 32] = True
    rr, cc = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    assert covered[cc < 3 * (rr // 3)].all()
    # the scratch layout of the substitution: rows back to back, 3 (r // 3) entries each
    assert row_off(0) == 0 and row_off(N - 1) + 3 * (B - 1) == KINV
    for r in range(N - 1):
        assert row_off(r + 1) == row_off(r) + 3 * (r // 3)
    # tensor rows parked at the top are taken out before the rows of X (growing from the bottom) reach them
    assert t_off(B - 1) + 6 * (B - 1) == KINV
    for i in range(1, B - 1):
        assert t_off(i + 1) == t_off(i) + 6 * i
        assert row_off(3 * i) + 9 * i <= t_off(i + 1)          # X rows of site i end below tensor row i + 1


def _random_block(rs, n):
    """Symmetric dipole tensor with zero 3x3 diagonal blocks for n sites, polarizabilities, fields, old dipoles."""
    pos = rs.uniform(0, 12.0, size=(n, 3))
    T = np.zeros((3 * n, 3 * n))
    for a in range(n):
        for b in range(a + 1, n):
            d = pos[a] - pos[b]
            r = np.linalg.norm(d)
            t = np.eye(3) / r**3 - 3.0 * np.outer(d, d) / r**5
            T[3 * a:3 * a + 3, 3 * b:3 * b + 3] = t
            T[3 * b:3 * b + 3, 3 * a:3 * a + 3] = t
    alpha = np.repeat(rs.uniform(0.2, 1.5, size=n), 3)
    return T, alpha, rs.normal(size=3 * n), 0.1 * rs.normal(size=3 * n)


def test_block_walk_is_a_triangular_solve():
    rs = np.random.RandomState(5)
    n = 24
    T, alpha, es, mu_old = _random_block(rs, n)
    outside = 0.3 * rs.normal(size=3 * n)                      # what the rest of the system contributes to acc
    acc0 = T @ mu_old + outside                                # running contraction when the block starts
    # reference semantics: site by site, new dipoles seen by the later sites
    mu = mu_old.copy()
    efi_ref = np.zeros(3 * n)
    for k in range(n):
        s = slice(3 * k, 3 * k + 3)
        acc_k = T[s] @ mu + outside[s]
        efi_ref[s] = -acc_k
        mu[s] = alpha[s] * (es[s] - acc_k)
    # the pipeline's form
    L = np.eye(3 * n) + alpha[:, None] * np.tril(T, -1)
    for k in range(n):                                          # same-site components do not couple (T_kk = 0)
        assert not L[3 * k:3 * k + 3, 3 * k:3 * k + 3].any() or np.allclose(L[3 * k:3 * k + 3, 3 * k:3 * k + 3], np.eye(3))
    rhs = alpha * es - mu_old - alpha * acc0
    dmu = np.linalg.solve(L, rhs)
    mu_new = mu_old + dmu
    assert np.abs(mu_new - mu).max() < 1e-13 * np.abs(mu).max()
    efi = mu_new / alpha - es                                   # k_gs_efi
    assert np.abs(efi - efi_ref).max() < 1e-12 * np.abs(efi_ref).max()
    # and the rows' contraction after the panel has been pushed into them equals the contraction of the new dipoles
    assert np.abs((acc0 + T @ dmu) - (T @ mu_new + outside)).max() < 1e-12


def test_forward_substitution_by_row_sites_gives_the_inverse():
    """k_gs_inverse: X[i][c] = -sum_{jc <= j < i} L_ij X[j][c] with X[jc][c] = e_q, by row sites."""
    rs = np.random.RandomState(9)
    n = 12
    T, alpha, _, _ = _random_block(rs, n)
    A = alpha[:, None] * np.tril(T, -1)
    X = np.zeros((3 * n, 3 * n))
    for i in range(1, n):
        for c in range(3 * i):
            jc, q = divmod(c, 3)
            x = -A[3 * i:3 * i + 3, 3 * jc + q].copy()
            for j in range(jc + 1, i):
                x -= A[3 * i:3 * i + 3, 3 * j:3 * j + 3] @ X[3 * j:3 * j + 3, c]
            X[3 * i:3 * i + 3, c] = x
    inv = np.linalg.inv(np.eye(3 * n) + A)
    assert np.abs((np.eye(3 * n) + X) - inv).max() < 1e-13
