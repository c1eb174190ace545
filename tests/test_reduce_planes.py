"""Bucket reduction by bit planes (msm_reduce_rows_kernel / msm_planes_step_kernel + the host finish in msm_host_phase):
the index arithmetic of the device kernels restated on integers (any abelian group will do) against sum_v v * B_v."""
import random

import pytest


def reduce_rows(buckets, K):
    per = len(buckets) // K
    rows = [0] * (2 * per)
    for t in range(per):
        run = s = 0
        for k in reversed(range(K)):
            run += buckets[t * K + k]
            s += run
        rows[t] = run
        rows[per + t] = s
    return rows, per


def planes_step(inp, rows, length):
    half = length // 2
    out = [0] * ((rows + 1) * half)
    for t in range((rows + 1) * half):
        r, i = divmod(t, half)
        out[t] = inp[2 * i + 1] if r == rows else inp[r * length + 2 * i] + inp[r * length + 2 * i + 1]
    return out


def host_finish(win, planes, K):
    result = win[0]
    if planes > 0:
        acc = win[planes]
        for b in range(planes - 2, -1, -1):
            acc = 2 * acc + win[1 + b]
        k = K
        while k > 1:
            acc *= 2
            k >>= 1
        result += acc
    return result


@pytest.mark.parametrize("log_nb,K", [(1, 2), (4, 16), (7, 2), (7, 4), (10, 16), (12, 64), (5, 32)])
def test_bit_plane_reduction_equals_weighted_bucket_sum(log_nb, K):
    rng = random.Random(log_nb * 100 + K)
    nb = 1 << log_nb
    buckets = [rng.randrange(1 << 64) if rng.random() < 0.8 else 0 for _ in range(nb)]
    cur, per = reduce_rows(buckets, K)
    rows, length = 2, per
    while length > 1:
        cur = planes_step(cur, rows, length)
        rows += 1
        length //= 2
    win = cur[1:rows]
    assert host_finish(win, rows - 2, K) == sum((i + 1) * b for i, b in enumerate(buckets))
