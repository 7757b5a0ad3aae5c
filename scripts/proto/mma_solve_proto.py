"""Matrix-level prototype of the DMMA block-tridiagonal solve (pk_solve_mma.cuh): validates the blocked right-looking
panel scheme (panel width 4), the rhs riding in padding row b of the diagonal tiles, the A' trick that carries the rhs
coupling into the next block, identity padding, and the two-chain (twisted) elimination meeting at the middle block."""
import numpy as np

def potrf4_inv(A44):
    L = np.linalg.cholesky(A44)
    return L, np.linalg.inv(L)

def chain_step(D, C, b, P):
    """D: PxP (sym full, row b = rhs^T in cols < b, zeros elsewhere in padding), C: PxP (rows next block, cols this block; padding 0).
    returns L (P x P lower, incl. row b = y^T), Z (P x P), Dn (P x P Schur accumulation incl. rhs row b)."""
    D = D.copy(); C = C.copy() if C is not None else None
    Dn = np.zeros((P, P))
    F_D = np.zeros((P, P)); F_C = np.zeros((P, P))
    for p in range(P // 4):
        c0 = 4 * p
        A44 = D[c0:c0 + 4, c0:c0 + 4].copy()
        for k in range(4):
            if c0 + k >= b:
                A44[k, :] = 0; A44[:, k] = 0; A44[k, k] = 1
        L44, W = potrf4_inv(A44)
        # panel solve for every row group (8 rows) at or below the panel's tile row
        fD = D[:, c0:c0 + 4] @ W.T
        # mask: rows above the panel = 0; rows of the panel: lower triangle only
        for r in range(P):
            if r < c0: fD[r, :] = 0
            elif r < c0 + 4: fD[r, r - c0 + 1:] = 0
        # tile rows above the panel's tile row are not computed at all
        fD[:8 * (p // 2), :] = 0
        F_D[:, c0:c0 + 4] = fD
        if C is not None:
            fC = C[:, c0:c0 + 4] @ W.T
            F_C[:, c0:c0 + 4] = fC
        # trailing updates (full tiles as the DMMAs do them; only tile columns that still hold unfactored columns)
        J = p // 2; h = p % 2
        for Jt in range(J if h == 0 else J + 1, P // 8):
            cols = slice(8 * Jt, 8 * Jt + 8)
            Bf = fD[cols, :].copy()            # f of D row group Jt as the B operand:
            for r in range(8):                 # the rhs row and the padding rows must not act as columns
                if 8 * Jt + r >= b: Bf[r, :] = 0
            for I in range(Jt, P // 8):
                rows = slice(8 * I, 8 * I + 8)
                D[rows, cols] -= fD[rows, :] @ Bf.T
            if Jt > J:                           # keep the diagonal tile symmetric: mirror update of the upper tile not needed (lower tiles only)
                pass
            if C is not None:
                C[:, cols] -= fC @ Bf.T
        if C is not None:
            A = fC.copy()
            A[b, :] = fD[b, :]                   # A' : the rhs row of the D tile row rides into the next block's rhs row
            for I in range(P // 8):
                for Jn in range(I + 1):
                    Dn[8 * I:8 * I + 8, 8 * Jn:8 * Jn + 8] -= A[8 * I:8 * I + 8, :] @ fC[8 * Jn:8 * Jn + 8, :].T
    return F_D, F_C, Dn

def solve(Hd, Ho, g, lam):
    """Hd[i]: b x b, Ho[i] = H_{i,i+1}: b x b, g[i]: b.  Solves (H + lam I) x = -g with the two-chain scheme."""
    N, b = len(Hd), Hd[0].shape[0]
    P = 8 * ((b + 1 + 7) // 8)
    m = N // 2
    def load_D(i):
        D = np.zeros((P, P)); D[:b, :b] = Hd[i] + lam * np.eye(b); D[b, :b] = -g[i]
        return D
    def load_C(i, down):
        C = np.zeros((P, P))
        C[:b, :b] = Ho[i].T if down else Ho[i - 1]     # rows: next block in sweep direction, cols: this block
        return C
    L = [None] * N; Z = [None] * N; y = [None] * N
    accA = np.zeros((P, P))
    for i in range(0, m):
        D = load_D(i) + accA
        F_D, F_C, accA = chain_step(D, load_C(i, True), b, P)
        L[i] = F_D[:b, :b]; y[i] = F_D[b, :b]; Z[i] = F_C[:b, :b]
    accB = np.zeros((P, P))
    for i in range(N - 1, m, -1):
        D = load_D(i) + accB
        F_D, F_C, accB = chain_step(D, load_C(i, False), b, P)
        L[i] = F_D[:b, :b]; y[i] = F_D[b, :b]; Z[i] = F_C[:b, :b]
    # the Dn tiles are only accumulated for the lower tiles: symmetrize the diagonal tiles? (they are full-tile DMMAs: already symmetric)
    D = load_D(m) + accA + accB
    # lower tiles only were accumulated for off-diagonal tiles: mirror
    for I in range(P // 8):
        for J in range(I):
            D[8 * J:8 * J + 8, 8 * I:8 * I + 8] = D[8 * I:8 * I + 8, 8 * J:8 * J + 8].T
    F_D, _, _ = chain_step(D, None, b, P)
    L[m] = F_D[:b, :b]; y[m] = F_D[b, :b]
    x = [None] * N
    x[m] = np.linalg.solve(L[m].T, y[m])
    for i in range(m - 1, -1, -1):
        x[i] = np.linalg.solve(L[i].T, y[i] - Z[i].T @ x[i + 1])
    for i in range(m + 1, N):
        x[i] = np.linalg.solve(L[i].T, y[i] - Z[i].T @ x[i - 1])
    return np.concatenate(x)

def main():
    rng = np.random.default_rng(0)
    for (N, b) in [(11, 14), (11, 6), (5, 4), (2, 14), (3, 10), (1, 14), (11, 8), (4, 2), (7, 12)]:
        n = N * b
        # random SPD block tridiagonal
        A = np.zeros((n, n))
        J = rng.standard_normal((3 * n, n))
        full = J.T @ J
        for i in range(N):
            A[i * b:(i + 1) * b, i * b:(i + 1) * b] = full[i * b:(i + 1) * b, i * b:(i + 1) * b] + 5 * np.eye(b)
            if i + 1 < N:
                blk = 0.3 * full[i * b:(i + 1) * b, (i + 1) * b:(i + 2) * b]
                A[i * b:(i + 1) * b, (i + 1) * b:(i + 2) * b] = blk
                A[(i + 1) * b:(i + 2) * b, i * b:(i + 1) * b] = blk.T
        A += n * 0.5 * np.eye(n)
        g = rng.standard_normal(n)
        lam = 0.7
        Hd = [A[i * b:(i + 1) * b, i * b:(i + 1) * b] for i in range(N)]
        Ho = [A[i * b:(i + 1) * b, (i + 1) * b:(i + 2) * b] for i in range(N - 1)]
        x = solve(Hd, Ho, [g[i * b:(i + 1) * b] for i in range(N)], lam)
        ref = np.linalg.solve(A + lam * np.eye(n), -g)
        print(N, b, "max rel err %.2e" % (np.abs(x - ref).max() / np.abs(ref).max()))

main()
