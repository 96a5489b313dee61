"""Independent numpy fp64 restatement of SURVEY.md appendix A, used only to pin oracle/nmf_oracle.cpp.

Written from the algorithm statements (not from the C++ oracle): BLAS matmuls and numpy solves instead
of hand loops and Householder QR, so an error in one is unlikely to be mirrored in the other.
"""
import numpy as np


def _normalize(W):
    s = (W * W).sum(axis=0)
    nz = s > 0
    W[:, nz] = W[:, nz] / np.sqrt(s[nz])
    return W


def _constraint(k, offdiag, diag):
    C = np.full((k, k), offdiag, dtype=np.float64)
    C[np.diag_indices(k)] = diag
    return C


def run(algorithm, V, W0, H0, iterations, eps, params=None, use_constant_w=False):
    params = params or {}
    V = np.asarray(V, dtype=np.float64)
    W = np.array(W0, dtype=np.float64)
    H = np.array(H0, dtype=np.float64)
    m, n = V.shape
    k = W.shape[1]
    vtv = (V * V).sum()
    checks = []
    if algorithm == "nsnmf":
        th = params["theta"]
        S = _constraint(k, th / k, (1.0 - th) + th / k)
    if algorithm == "ahcls":
        bW = ((1 - params["alphaW"]) * np.sqrt(k) + params["alphaW"]) ** 2
        bH = ((1 - params["alphaH"]) * np.sqrt(k) + params["alphaH"]) ** 2
    for it in range(1, iterations + 1):
        check = it % 10 == 0 or it == iterations
        if algorithm == "mu":
            A = W.T @ W
            N = W.T @ V
            H = H * N / (A @ H + eps)
            if check:
                t2 = (H * N).sum()
                t3 = np.trace((H @ H.T) @ A)
            if not use_constant_w:
                B = H @ H.T
                W = _normalize(W * (V @ H.T) / (W @ B + eps))
        elif algorithm == "nsnmf":
            Wt = W @ S
            A = Wt.T @ Wt
            N = Wt.T @ V
            H = H * N / (A @ H + eps)
            Ht = S @ H
            B = Ht @ Ht.T
            if check:
                t2 = (H * N).sum()
                t3 = np.trace(B @ (W.T @ W))
            if not use_constant_w:
                W = _normalize(W * (V @ Ht.T) / (W @ B + eps))
        else:
            A = W.T @ W
            if algorithm == "gdcls":
                G = A + _constraint(k, 0.0, params["lambda"])
            elif algorithm == "als":
                G = A.copy()
            elif algorithm == "acls":
                G = A + _constraint(k, 0.0, params["lambdaH"])
            else:
                G = A + _constraint(k, -params["lambdaH"], params["lambdaH"] * bH - params["lambdaH"])
            H = np.maximum(0.0, np.linalg.solve(G, W.T @ V))
            B = H @ H.T
            if check:
                t3 = np.trace(B @ A)
            P = V @ H.T
            if algorithm == "gdcls":
                if not use_constant_w:
                    W = _normalize(W * P / (W @ B + eps))
                if check:
                    t2 = (P * W).sum()
            else:
                if check:
                    t2 = (W * P).sum()
                if not use_constant_w:
                    if algorithm == "acls":
                        B = B + _constraint(k, 0.0, params["lambdaW"])
                    elif algorithm == "ahcls":
                        B = B + _constraint(k, -params["lambdaW"], params["lambdaW"] * bW - params["lambdaW"])
                    W = _normalize(np.maximum(0.0, np.linalg.solve(B.T, P.T).T))
        if check:
            Wout = W @ S if algorithm == "nsnmf" else W
            checks.append((it, np.sqrt(vtv - 2 * t2 + t3), np.linalg.norm(V - Wout @ H)))
    if algorithm == "nsnmf":
        W = W @ S
    return W, H, checks
