"""N2 throughput: reference-style row-wise batch assembly + negative-sampling loop (oracle/ref_reader.py, the
restatement pinned by the reference-run golden file) against the tensor-native reader, on the host cores
(and with the tables resident in HBM when a GPU is present).  Prints one JSON line.

    python tools/bench_reader.py [--rows 1000000] [--batch 4096]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_reader  # noqa: E402  (tools/ is measurement infrastructure, like bench.py's cpu_baseline leg)
from pytorchrec_b200.data import TensorDataReader  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--users", type=int, default=100_000)
    ap.add_argument("--items", type=int, default=50_000)
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    n = a.rows
    train = {"uid": rng.integers(1, a.users + 1, n).astype(np.int32), "iid": rng.integers(1, a.items + 1, n).astype(np.int32),
             "rate": rng.integers(1, 6, n).astype(np.int32), "label": np.ones(n, dtype=np.int32),
             "time": np.arange(n, dtype=np.int32)}
    items = {"iid": np.arange(1, a.items + 1, dtype=np.int32), "i_c_cat": rng.integers(1, 100, a.items).astype(np.int32)}
    pos = {}
    for u, i in zip(train["uid"].tolist(), train["iid"].tolist()):
        pos.setdefault(u, set()).add(i)
    out = {"rows": n, "batch": a.batch, "cores": os.cpu_count()}

    # reference idiom, bounded sample
    t = time.perf_counter()
    neg = ref_reader.train_neg_sample_ref(np.random.default_rng(1), train["uid"][:200_000].tolist(), pos, 1, a.items + 1)
    out["rowwise_neg_sample_rows_per_s"] = 200_000 / (time.perf_counter() - t)
    pairs = np.stack([train["iid"][:200_000], neg], axis=1)
    t = time.perf_counter()
    nb = 8
    for b in range(nb):
        ref_reader.assemble_batch_ref(train, np.arange(b * a.batch, (b + 1) * a.batch) % 200_000, items, pairs)
    out["rowwise_assembly_samples_per_s"] = nb * a.batch / (time.perf_counter() - t)

    devices = [torch.device("cpu")] + ([torch.device("cuda:0")] if torch.cuda.is_available() else [])
    for dev in devices:
        for sampler in ("reference", "device"):
            r = TensorDataReader(train, items=items, train_mode="pair_wise", user_pos_his_set_dict=pos,
                                 rng=np.random.default_rng(1), device=dev, sampler=sampler)
            r.train_neg_sample()
            if dev.type == "cuda":
                torch.cuda.synchronize()
            t = time.perf_counter()
            r.train_neg_sample()
            if dev.type == "cuda":
                torch.cuda.synchronize()
            out[f"tensor_neg_sample_{sampler}_{dev.type}_rows_per_s"] = len(r._pair_np) / (time.perf_counter() - t)
        t = time.perf_counter()
        cnt = 0
        for batch in r.batches("train", a.batch, shuffle=True):
            cnt += batch["uid"].shape[0]
        if dev.type == "cuda":
            torch.cuda.synchronize()
        out[f"tensor_assembly_{dev.type}_samples_per_s"] = cnt / (time.perf_counter() - t)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
