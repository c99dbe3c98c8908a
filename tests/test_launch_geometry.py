"""The lane -> (piece, node) mapping of k_penalty (topay_b200/csrc/solver_kernels.cuh, solver.cu::launch_eval),
restated in Python: every penalty node of every piece must be evaluated exactly once, whatever the rounding of
the grid to whole blocks does. (A spare warp once re-evaluated node 2K when K < K-hat.)"""
import pytest

PEN_WARPS = 2          # TP_PEN_WARPS


def kpad_of(K):
    p = 4
    while p < K:
        p <<= 1
    return p


def coverage(K, max_N, N):
    Kpad = kpad_of(K)
    ppw = 32 // Kpad
    end_tasks = K == Kpad
    groups = (max_N + ppw - 1) // ppw
    end_warps = (max_N + 31) // 32 if end_tasks else 0
    blocks = (groups + end_warps + PEN_WARPS - 1) // PEN_WARPS
    seen = {}
    for task in range(blocks * PEN_WARPS):
        for lane in range(32):
            if task >= groups:
                if not end_tasks:
                    continue                      # the spare warp returns
                piece = (task - groups) * 32 + lane
                if piece < N:
                    seen[(piece, K)] = seen.get((piece, K), 0) + 1
            else:
                seg, jn = divmod(lane, Kpad)
                piece = task * ppw + seg
                n_nodes = K if end_tasks else K + 1
                if piece < N and jn < n_nodes:
                    seen[(piece, jn)] = seen.get((piece, jn), 0) + 1
    return seen


@pytest.mark.parametrize("K", [1, 3, 4, 5, 8, 12, 16, 31, 32])
def test_every_node_exactly_once(K):
    for max_N in (1, 2, 3, 5, 6, 7, 16, 33, 64):
        for N in {1, max(1, max_N // 2), max_N}:
            seen = coverage(K, max_N, N)
            assert set(seen) == {(p, j) for p in range(N) for j in range(K + 1)}, (K, max_N, N)
            assert all(v == 1 for v in seen.values()), (K, max_N, N)
