"""CPU model of the exchange pattern of attn_cluster_push_kernel (csrc/attention.cuh): every CTA
of the cluster pushes slices of its partial into the inbox of the CTA that merges them; after one
cluster barrier each CTA merges its items from local memory.  Checks, for the shapes the engine
uses, that every (row, 16-dim segment) item is merged exactly once, from the partials of ALL
splits in split order, and that the inbox never overflows the shared memory the host reserves."""
import numpy as np
import pytest

THREADS, STRIDE = 128, 20


def smem_inbox_floats(rows_pad):            # attn_push_smem_bytes: (rows_pad * 8 + 8) entries
    return (rows_pad * 8 + 8) * STRIDE


@pytest.mark.parametrize("group,M,n_splits", [(1, 1, 8), (1, 7, 8), (1, 16, 8), (4, 7, 8), (4, 16, 8),
                                              (8, 7, 8), (8, 16, 8), (1, 7, 3), (4, 9, 5), (2, 16, 1)])
def test_every_item_is_merged_once_from_all_splits_in_order(group, M, n_splits):
    R = group * M
    rows_pad = group * 16
    n_items = R * 8
    cap = (n_items + n_splits - 1) // n_splits
    assert n_splits * cap * STRIDE <= smem_inbox_floats(rows_pad)
    # partial of split s, row r, segment d: a unique tag; (m, l) tagged likewise
    inbox = [np.full(smem_inbox_floats(rows_pad), -1.0) for _ in range(n_splits)]
    for split in range(n_splits):                          # push phase of CTA `split`
        for tid in range(THREADS):
            for item in range(tid, n_items, THREADS):
                row, seg = item >> 3, item & 7
                dest, li = item % n_splits, item // n_splits
                base = (split * cap + li) * STRIDE
                assert inbox[dest][base] == -1.0, "slot written twice"
                inbox[dest][base:base + 16] = split * 1e6 + row * 1e3 + seg
                inbox[dest][base + 16:base + 18] = [split * 1e6 + row * 1e3 + 900, split * 1e6 + row * 1e3 + 901]
    merged = {}
    for split in range(n_splits):                          # merge phase of CTA `split`
        for tid in range(THREADS):
            for li in range(tid, cap, THREADS):
                item = li * n_splits + split
                if item >= n_items:
                    break
                row, seg = item >> 3, item & 7
                srcs = []
                for s in range(n_splits):
                    base = (s * cap + li) * STRIDE
                    assert inbox[split][base] == s * 1e6 + row * 1e3 + seg
                    assert inbox[split][base + 16] == s * 1e6 + row * 1e3 + 900      # m of (split s, row)
                    assert inbox[split][base + 17] == s * 1e6 + row * 1e3 + 901      # l
                    srcs.append(s)
                assert (row, seg) not in merged
                merged[(row, seg)] = srcs
    assert len(merged) == n_items
    assert all(v == list(range(n_splits)) for v in merged.values())
