"""Property tests (hypothesis) for the integer / index work of the path, as SURVEY.md §4 plans: randomly drawn shapes,
masks and id distributions instead of hand-picked cases.

* CPU (``-m "not gpu"``): invariants of the oracle restatements themselves — they are the checker, so their own
  consistency is worth a test: offsets are the cumulative valid counts, the sorted keys are a stable permutation of the
  masked keys, ``unique / counts`` tile the sorted list, the mixed-radix cross is injective.
* GPU (``-m gpu``): ``ptrec_index_prep`` and ``ptrec_sort_dedup`` (all three sort paths) against the oracle, bit for
  bit, and the one-hot gather against ``F.embedding``."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import ref_ops

MASKS = ["none", "pad", "pad_keep_first", "lens"]
SETTINGS = dict(deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])


def _draw_ids(rng, shape, rows, pad_frac, hot):
    x = rng.integers(0, rows, size=shape)
    if hot:  # a few hot ids: long duplicate runs
        x = np.where(rng.random(shape) < 0.5, rng.integers(0, min(rows, 3), size=shape), x)
    if pad_frac > 0:
        x[rng.random(shape) < pad_frac] = 0
    return torch.from_numpy(x.astype(np.int64))


@settings(max_examples=60, **SETTINGS)
@given(B=st.integers(1, 300), L=st.integers(1, 40), mask=st.sampled_from(MASKS), seed=st.integers(0, 2 ** 20),
       pad_frac=st.sampled_from([0.0, 0.3, 0.9, 1.0]))
def test_oracle_index_prep_invariants(B, L, mask, seed, pad_frac):
    rng = np.random.default_rng(seed)
    ids = _draw_ids(rng, (B, L), 50, pad_frac, False)
    lens = torch.from_numpy(rng.integers(0, L + 1, size=B).astype(np.int32))
    out, off = ref_ops.index_prep_ref(ids, mask, lens)
    valid = ref_ops.valid_mask(ids, mask, lens)
    assert off[0] == 0 and torch.equal(off[1:] - off[:-1], valid.sum(1)) and int(off[-1]) == out.numel()
    for b in rng.integers(0, B, size=min(B, 5)):          # each bag's slice = its valid ids in slot order
        assert torch.equal(out[off[b]:off[b + 1]], ids[b][valid[b]])
    if mask == "pad_keep_first":
        assert bool(valid[:, 0].all())
    if mask == "none":
        assert out.numel() == B * L


@settings(max_examples=60, **SETTINGS)
@given(n=st.integers(1, 3000), rows=st.sampled_from([1, 2, 7, 300, 70000, 1 << 20]), seed=st.integers(0, 2 ** 20),
       hot=st.booleans(), bad_frac=st.sampled_from([0.0, 0.1]))
def test_oracle_sort_dedup_invariants(n, rows, seed, hot, bad_frac):
    rng = np.random.default_rng(seed)
    ids = _draw_ids(rng, (n,), rows, 0.0, hot)
    ids[rng.random(n) < bad_frac] = rows + 5            # out of range -> masked
    valid = torch.from_numpy(rng.random(n) < 0.9)
    (skey, perm, uniq, counts), = ref_ops.sort_dedup_ref([ids], [valid], [rows])
    key = ids.clone()
    key[~(valid & (ids >= 0) & (ids < rows))] = ref_ops.MASKED
    assert torch.equal(key[perm], skey) and bool((skey[1:] >= skey[:-1]).all())
    assert torch.equal(torch.sort(perm).values, torch.arange(n))
    same = skey[1:] == skey[:-1]                        # stability: equal keys keep slot order
    assert bool((perm[1:][same] > perm[:-1][same]).all())
    assert int(counts.sum()) == n and torch.equal(torch.repeat_interleave(uniq, counts), skey)


@settings(max_examples=40, **SETTINGS)
@given(cards=st.lists(st.integers(1, 50), min_size=1, max_size=4), n=st.integers(1, 200), seed=st.integers(0, 2 ** 20))
def test_oracle_crossed_ids_are_a_bijection_of_the_tuple(cards, n, seed):
    rng = np.random.default_rng(seed)
    cols = [rng.integers(0, c, size=n) for c in cards]
    x = ref_ops.crossed_ids_ref(cols, cards)
    assert x.min() >= 0 and x.max() < int(np.prod(cards))
    back = []
    y = x.copy()
    for c in reversed(cards):
        back.append(y % c)
        y = y // c
    for got, want in zip(reversed(back), cols):
        assert np.array_equal(got, want)


# ----------------------------------------------------------------------------------------------------- GPU
gpu = pytest.mark.gpu


@gpu
@settings(max_examples=40, **SETTINGS)
@given(B=st.integers(1, 5000), L=st.integers(1, 120), mask=st.sampled_from(MASKS), seed=st.integers(0, 2 ** 20),
       pad_frac=st.sampled_from([0.0, 0.5, 1.0]))
def test_index_prep_matches_the_oracle(B, L, mask, seed, pad_frac):
    from pytorchrec_b200 import ops
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(seed)
    ids = _draw_ids(rng, (B, L), 1000, pad_frac, False)
    lens = torch.from_numpy(rng.integers(0, L + 1, size=B).astype(np.int32))
    out, off = ops.index_prep(ids.to(dev), lens.to(dev) if mask == "lens" else None, mask)
    ref_ids, ref_off = ref_ops.index_prep_ref(ids, mask, lens)
    assert torch.equal(off.cpu(), ref_off)
    assert torch.equal(out.cpu()[: int(ref_off[-1])], ref_ids)


@gpu
@pytest.mark.parametrize("path", ["radix", "smem", "one_sweep"])
@settings(max_examples=25, **SETTINGS)
@given(B=st.integers(1, 9000), rows=st.lists(st.sampled_from([1, 3, 257, 60000, 1 << 20, 1 << 26]), min_size=1, max_size=3),
       seed=st.integers(0, 2 ** 20), hot=st.booleans())
def test_sort_dedup_matches_the_oracle(path, B, rows, seed, hot):
    from pytorchrec_b200 import _lib, ops
    dev = torch.device("cuda:0")
    lib = _lib.load()
    before = lib.ptrec_one_sweep_sort_enabled()
    lib.ptrec_set_smem_sort(2 if path == "smem" else 0)
    lib.ptrec_set_one_sweep_sort(1 if path == "one_sweep" else 0)
    try:
        rng = np.random.default_rng(seed)
        T = len(rows)
        id_list = [_draw_ids(rng, (B,), rows[t], 0.0, hot) for t in range(T)]
        for t in range(T):                               # a few out-of-range and negative ids: masked keys
            bad = rng.random(B) < 0.02
            id_list[t][torch.from_numpy(bad)] = rows[t] + 1
        layout = ops.FeatureLayout([dict(table=t, bag_len=1) for t in range(T)], 4, T)
        tables = ops.TableSet()
        tables.ptrs = torch.zeros(T, dtype=torch.int64, device=dev)
        tables.rows = torch.tensor(rows, dtype=torch.int64, device=dev)
        tables.max_rows = max(rows)
        srt = ops.sort_dedup(tables, layout, torch.cat(id_list).to(dev), None, B)
        torch.cuda.synchronize()
        keys = srt.sorted_keys.cpu()[: srt.N].long() & 0xFFFFFFFF
        perm = srt.perm.cpu()[: srt.N].long()
        ref = ref_ops.sort_dedup_ref(id_list, [torch.ones(B, dtype=torch.bool)] * T, rows)
        n_seg, pos = 0, 0
        for t, (skey, rperm, uniq, counts) in enumerate(ref):
            assert torch.equal(keys[pos:pos + B], skey & 0xFFFFFFFF), (path, t)
            assert torch.equal(perm[pos:pos + B], rperm + pos), (path, t)
            n_seg += int(uniq.numel())   # the masked key forms a segment of its own (the update skips it)
            pos += B
        assert int(srt.n_seg.item()) == n_seg
    finally:
        lib.ptrec_set_smem_sort(1)
        lib.ptrec_set_one_sweep_sort(before)


@gpu
@settings(max_examples=25, **SETTINGS)
@given(B=st.integers(1, 3000), D=st.sampled_from([1, 2, 4, 8, 16, 32, 64, 128]), T=st.integers(1, 5),
       seed=st.integers(0, 2 ** 20))
def test_onehot_gather_copies_rows_bit_for_bit(B, D, T, seed):
    from pytorchrec_b200 import ops
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(seed)
    rows = [int(r) for r in rng.integers(1, 2000, size=T)]
    g = torch.Generator().manual_seed(seed)
    weights = [torch.randn(r, D, generator=g) for r in rows]
    id_list = [_draw_ids(rng, (B,), rows[t], 0.0, False) for t in range(T)]
    layout = ops.FeatureLayout([dict(table=t, bag_len=1) for t in range(T)], D, T)
    dw = [w.to(dev) for w in weights]   # kept alive: the table set holds raw pointers
    tables = ops.TableSet().refresh(dw)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    out, _ = ops.gather_pool_fwd(tables, layout, torch.cat(id_list).to(dev), None, B, err_flag=err)
    ref = torch.stack([torch.nn.functional.embedding(id_list[t], weights[t]) for t in range(T)], 1)
    assert torch.equal(out.view(B, T, D).cpu(), ref) and err.item() == 0
