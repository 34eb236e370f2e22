"""Kernel-level parity on the GPU: every call goes through the C ABI (ctypes) and is compared with
the CPU oracle on the same seeded inputs.  Integer artefacts are bit-exact; fp32 outputs within
1e-5 relative (segment-scaled where accumulation order differs)."""
import numpy as np
import pytest
import torch

from oracle import ref_ops
from pytorchrec_b200 import _lib, ops

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
RTOL = 1e-5


def _tables(rows, D, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(r, D, generator=g) for r in rows]


def _ids(shape, rows, seed, zipf=False, pad_frac=0.0):
    rng = np.random.default_rng(seed)
    if zipf:
        x = rng.zipf(1.3, size=shape) % rows
    else:
        x = rng.integers(0, rows, size=shape)
    if pad_frac > 0:
        x[rng.random(shape) < pad_frac] = 0
    return torch.from_numpy(x.astype(np.int64))


def _u32(t):
    return t.long() & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------ index prep
@pytest.mark.parametrize("mask", ["none", "pad", "pad_keep_first", "lens"])
@pytest.mark.parametrize("B,L", [(1, 1), (7, 5), (300, 33), (4099, 100)])
def test_index_prep_bit_exact(mask, B, L):
    ids = _ids((B, L), 50, seed=B * 131 + L, pad_frac=0.5)
    lens = torch.from_numpy(np.random.default_rng(B).integers(0, L + 1, size=B).astype(np.int32))
    out_ids, offsets = ops.index_prep(ids.to(DEV), lens.to(DEV) if mask == "lens" else None, mask)
    ref_ids, ref_off = ref_ops.index_prep_ref(ids, mask, lens)
    assert torch.equal(offsets.cpu(), ref_off)
    n = int(ref_off[-1])
    assert torch.equal(out_ids.cpu()[:n], ref_ids)


def test_index_prep_all_pad_rows_are_empty_bags():
    ids = torch.zeros(9, 4, dtype=torch.int64)
    ids[3, 1] = 5
    out_ids, offsets = ops.index_prep(ids.to(DEV), None, "pad")
    assert offsets.cpu().tolist() == [0, 0, 0, 0, 1, 1, 1, 1, 1, 1]
    assert out_ids.cpu()[0].item() == 5


# ------------------------------------------------------------------------------------------ gather + pool
def _run_gather(weights, specs, id_list, lens_list, D):
    """specs: per feature dict(table, pooling, mask); returns CUDA out [B, F, D] and oracle out."""
    F = len(specs)
    B = id_list[0].shape[0]
    lay_specs, lens_cols = [], []
    for f, s in enumerate(specs):
        L = 1 if id_list[f].dim() == 1 else id_list[f].shape[1]
        lc = -1
        if s.get("mask") == "lens":
            lens_cols.append(lens_list[f])
            lc = len(lens_cols) - 1
        lay_specs.append(dict(table=s["table"], bag_len=L, pooling=s.get("pooling", "sum"),
                              mask=s.get("mask", "none"), lens_col=lc))
    layout = ops.FeatureLayout(lay_specs, D, len(weights))
    dw = [w.to(DEV) for w in weights]
    tables = ops.TableSet().refresh(dw)
    ids = torch.cat([i.reshape(-1) for i in id_list]).to(DEV)
    lens = torch.stack(lens_cols).to(torch.int32).to(DEV) if lens_cols else None
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    out, scale = ops.gather_pool_fwd(tables, layout, ids, lens, B, want_scale=True, err_flag=err)
    ref = ref_ops.multi_table_lookup_ref(weights, [s["table"] for s in specs], id_list,
                                         [s.get("pooling", "sum") for s in specs],
                                         [s.get("mask", "none") for s in specs], lens_list)
    return out.view(B, F, D).cpu(), ref, err, (layout, tables, ids, lens, scale, dw)


@pytest.mark.parametrize("D", [1, 2, 4, 8, 12, 16, 32, 64, 128])
@pytest.mark.parametrize("B", [1, 33, 1024])
def test_onehot_gather_is_bit_exact(D, B):
    rows = [17, 1000, 5, 301]
    weights = _tables(rows, D, seed=D)
    id_list = [_ids((B,), rows[t], seed=100 * D + t) for t in range(4)]
    out, ref, err, _ = _run_gather(weights, [dict(table=t) for t in range(4)], id_list, [None] * 4, D)
    assert torch.equal(out, ref)  # values are copied, not computed
    assert err.item() == 0


def test_onehot_gather_criteo_shape_bit_exact():
    F, D, B = 26, 16, 4096
    rows = [1000 + 37 * f for f in range(F)]
    weights = _tables(rows, D, seed=1)
    id_list = [_ids((B,), rows[f], seed=f, zipf=(f % 2 == 0)) for f in range(F)]
    out, ref, err, _ = _run_gather(weights, [dict(table=f) for f in range(F)], id_list, [None] * F, D)
    assert torch.equal(out, ref) and err.item() == 0


@pytest.mark.parametrize("pooling", ["sum", "mean", "sqrtn"])
@pytest.mark.parametrize("mask", ["none", "pad", "pad_keep_first", "lens"])
@pytest.mark.parametrize("B,L,D", [(5, 3, 8), (130, 100, 16), (64, 7, 32), (33, 50, 64), (9, 11, 1)])
def test_bag_pooling_matches_reference_idioms(pooling, mask, B, L, D):
    rows = [211, 97]
    weights = _tables(rows, D, seed=L)
    id_list = [_ids((B, L), rows[0], seed=B + L, pad_frac=0.4), _ids((B,), rows[1], seed=3)]
    lens = torch.from_numpy(np.random.default_rng(L).integers(0, L + 1, size=B).astype(np.int64))
    specs = [dict(table=0, pooling=pooling, mask=mask), dict(table=1)]
    out, ref, err, _ = _run_gather(weights, specs, id_list, [lens, None], D)
    # segment-scaled tolerance: |x - ref| <= 1e-5 * sum_l |v_l| (accumulation order differs)
    bound = RTOL * torch.nn.functional.embedding(id_list[0], weights[0]).abs().sum(1) + 1e-7
    assert ((out[:, 0] - ref[:, 0]).abs() <= bound).all()
    assert torch.equal(out[:, 1], ref[:, 1])
    assert err.item() == 0


def test_shared_table_and_long_unstaged_bags():
    D, B, L = 16, 6, 2500  # L > staging capacity -> direct index reads
    weights = _tables([400], D, seed=9)
    id_list = [_ids((B,), 400, seed=1), _ids((B, L), 400, seed=2, pad_frac=0.2)]
    specs = [dict(table=0), dict(table=0, pooling="mean", mask="pad")]
    out, ref, err, _ = _run_gather(weights, specs, id_list, [None, None], D)
    assert torch.equal(out[:, 0], ref[:, 0])
    np.testing.assert_allclose(out[:, 1].numpy(), ref[:, 1].numpy(), rtol=1e-4, atol=1e-5)


def test_out_of_range_id_sets_error_flag_and_zero_row():
    weights = _tables([10], 8, seed=0)
    ids = torch.tensor([1, 10, 3, -1], dtype=torch.int64)
    layout = ops.FeatureLayout([dict(table=0, bag_len=1)], 8, 1)
    tables = ops.TableSet().refresh([weights[0].to(DEV)])
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    out, _ = ops.gather_pool_fwd(tables, layout, ids.to(DEV), None, 4, err_flag=err)
    assert err.item() == 1
    assert torch.equal(out.cpu()[[0, 2]], weights[0][[1, 3]])
    assert (out.cpu()[[1, 3]] == 0).all()


# ------------------------------------------------------------------------------------------ sort + dedup
def _check_sort(layout, tables, ids_dev, lens_dev, B, specs, id_list, lens_list, rows):
    srt = ops.sort_dedup(tables, layout, ids_dev, lens_dev, B)
    torch.cuda.synchronize()
    N = srt.N
    keys = _u32(srt.sorted_keys.cpu()[:N])
    perm = srt.perm.cpu()[:N].long()
    n_seg = int(srt.n_seg.item())
    seg_start = srt.seg_start.cpu()[: n_seg + 1].long()
    seg_table = srt.seg_table.cpu()[:n_seg].long()
    meta = srt.seg_meta.cpu()[:n_seg]
    assert torch.equal(_u32(meta[:, 0]), keys[seg_start[:n_seg]]), "segment record key"
    assert torch.equal(meta[:, 1].long(), perm[seg_start[:n_seg]]), "segment record first slot"
    # oracle: per table, concatenate its features' slots in layout order
    T = len(rows)
    per_ids, per_valid, base = [], [], []
    pos = 0
    order = sorted(range(len(specs)), key=lambda f: (specs[f]["table"], f))
    assert order == list(range(len(specs))), "test specs must already be ordered by table"
    for t in range(T):
        ids_t, val_t = [], []
        base.append(pos)
        for f, s in enumerate(specs):
            if s["table"] != t:
                continue
            x = id_list[f] if id_list[f].dim() == 2 else id_list[f].unsqueeze(1)
            v = ref_ops.valid_mask(x, s.get("mask", "none") if id_list[f].dim() == 2 or s.get("mask") == "lens" else "none",
                                   lens_list[f])
            ids_t.append(x.reshape(-1))
            val_t.append(v.reshape(-1))
            pos += x.numel()
        per_ids.append(torch.cat(ids_t) if ids_t else torch.zeros(0, dtype=torch.int64))
        per_valid.append(torch.cat(val_t) if val_t else torch.zeros(0, dtype=torch.bool))
    ref = ref_ops.sort_dedup_ref(per_ids, per_valid, rows)
    seg = 0
    for t in range(T):
        skey, rperm, uniq, counts = ref[t]
        n_t = skey.numel()
        assert torch.equal(keys[base[t]: base[t] + n_t], skey), f"sorted keys differ in table {t}"
        assert torch.equal(perm[base[t]: base[t] + n_t] - base[t], rperm), f"perm differs in table {t}"
        k = uniq.numel()
        assert torch.equal(keys[seg_start[seg: seg + k]], uniq), f"unique ids differ in table {t}"
        assert torch.equal(seg_start[seg + 1: seg + k + 1] - seg_start[seg: seg + k], counts)
        assert (seg_table[seg: seg + k] == t).all()
        seg += k
    assert seg == n_seg and int(seg_start[n_seg]) == N
    return srt


@pytest.fixture(params=["smem_sort", "global_sort", "one_sweep_sort"])
def sort_path(request):
    """K2a has a one-CTA-per-table shared-memory path, a multi-launch global path (histogram / scan / scatter per
    digit) and a one-sweep global path (decoupled look-back): same outputs from all three."""
    lib = _lib.load()
    before = lib.ptrec_one_sweep_sort_enabled()
    lib.ptrec_set_smem_sort(2 if request.param == "smem_sort" else 0)
    lib.ptrec_set_one_sweep_sort(1 if request.param == "one_sweep_sort" else 0)
    yield request.param
    lib.ptrec_set_smem_sort(1)
    lib.ptrec_set_one_sweep_sort(before)


@pytest.mark.parametrize("B", [1, 100, 2048, 5000, 22528, 30000])
@pytest.mark.parametrize("rows", [[3, 70000, 257], [1 << 20, 9, 300], [60000, 200, 7], [5, 1 << 26, 999]],
                         ids=["17bit", "20bit", "16bit_8bit_digits", "26bit_3_passes"])
def test_sort_dedup_bit_exact_onehot(B, rows, sort_path):
    D = 4
    weights = _tables([min(r, 64) for r in rows], D, seed=1)  # contents irrelevant: only row counts matter
    id_list = [_ids((B,), rows[t], seed=B + t, zipf=(t == 1)) for t in range(3)]
    specs = [dict(table=t) for t in range(3)]
    layout = ops.FeatureLayout([dict(table=t, bag_len=1) for t in range(3)], D, 3)
    tables = ops.TableSet()
    tables.ptrs = torch.zeros(3, dtype=torch.int64, device=DEV)
    tables.rows = torch.tensor(rows, dtype=torch.int64, device=DEV)
    tables.max_rows = max(rows)
    ids = torch.cat(id_list).to(DEV)
    _check_sort(layout, tables, ids, None, B, specs, id_list, [None] * 3, rows)


@pytest.mark.parametrize("mask", ["pad", "pad_keep_first", "lens"])
def test_sort_dedup_bags_masks_and_shared_table(mask, sort_path):
    B, L, D = 257, 19, 8
    rows = [40, 1000]
    id_list = [_ids((B,), rows[0], seed=1), _ids((B, L), rows[0], seed=2, pad_frac=0.5), _ids((B, 3), rows[1], seed=3)]
    lens = torch.from_numpy(np.random.default_rng(5).integers(0, L + 1, size=B).astype(np.int64))
    specs = [dict(table=0), dict(table=0, mask=mask), dict(table=1)]
    layout = ops.FeatureLayout([dict(table=0, bag_len=1), dict(table=0, bag_len=L, mask=mask, lens_col=0 if mask == "lens" else -1),
                                dict(table=1, bag_len=3)], D, 2)
    tables = ops.TableSet()
    tables.ptrs = torch.zeros(2, dtype=torch.int64, device=DEV)
    tables.rows = torch.tensor(rows, dtype=torch.int64, device=DEV)
    tables.max_rows = max(rows)
    ids = torch.cat([i.reshape(-1) for i in id_list]).to(DEV)
    lens_dev = lens.to(torch.int32).reshape(1, B).to(DEV) if mask == "lens" else None
    _check_sort(layout, tables, ids, lens_dev, B, specs, id_list, [None, lens, None], rows)


def test_sort_dedup_out_of_range_ids_are_masked(sort_path):
    rows = [10]
    ids = torch.tensor([3, 10, -4, 3, 0, 9], dtype=torch.int64)
    layout = ops.FeatureLayout([dict(table=0, bag_len=1)], 4, 1)
    tables = ops.TableSet()
    tables.ptrs = torch.zeros(1, dtype=torch.int64, device=DEV)
    tables.rows = torch.tensor(rows, dtype=torch.int64, device=DEV)
    tables.max_rows = 10
    srt = ops.sort_dedup(tables, layout, ids.to(DEV), None, 6)
    assert _u32(srt.sorted_keys.cpu()).tolist() == [0, 3, 3, 9, 0xFFFFFFFF, 0xFFFFFFFF]
    assert srt.perm.cpu().tolist() == [4, 0, 3, 5, 1, 2]
    assert srt.n_seg.item() == 4


# ------------------------------------------------------------------------------------------ fused update
def _dense_grads(weights, specs, id_list, lens_list, grad_out):
    """Reference backward: dense [rows, D] gradient per table via autograd on the oracle forward."""
    ws = [w.clone().requires_grad_(True) for w in weights]
    out = ref_ops.multi_table_lookup_ref(ws, [s["table"] for s in specs], id_list,
                                         [s.get("pooling", "sum") for s in specs],
                                         [s.get("mask", "none") for s in specs], lens_list)
    out.backward(grad_out)
    return [w.grad if w.grad is not None else torch.zeros_like(w) for w in ws]


OPT_CASES = [
    ("sgd", dict(lr=0.3)),
    ("adagrad", dict(lr=0.2, eps=1e-10, lr_decay=0.01)),
    ("rowwise_adagrad", dict(lr=0.2, eps=1e-8)),
    ("lazy_adam", dict(lr=0.05, beta1=0.9, beta2=0.999, eps=1e-8)),
]


@pytest.mark.parametrize("opt_name,hp", OPT_CASES)
@pytest.mark.parametrize("D", [1, 8, 16, 64])
@pytest.mark.parametrize("hot", [False, True])
def test_fused_update_matches_dense_optimizers(opt_name, hp, D, hot):
    """Three steps of sort+dedup+fused update vs torch's dense optimizer fed the dense reference gradient.
    `hot` makes ids Zipf-heavy so that runs longer than 32 exercise the CTA-per-segment kernel."""
    _fused_update_case(opt_name, hp, D, hot)


@pytest.mark.parametrize("D", [1, 16, 64])
def test_fused_update_giant_runs_are_split_over_ctas(D):
    """A row hit by a large share of the batch (the padding id of every history in a DIN batch; a Zipf head at a large
    batch) is one run of 1e4 - 1e5 slots: it is reduced chunk by chunk by many CTAs and the chunks are added in order.
    Checked against the fp64 sum of its gradient rows (Adagrad, one step) and for run-to-run bit-reproducibility."""
    B, rows = 50000, [300, 7]
    lay = ops.FeatureLayout([dict(table=0, bag_len=1), dict(table=1, bag_len=1)], D, 2)
    g = torch.Generator().manual_seed(D)
    rng = np.random.default_rng(D)
    ids0 = rng.integers(0, rows[0], size=B)
    ids0[rng.random(B) < 0.6] = 0              # ~30000 slots on row 0: 15 chunks
    ids1 = rng.integers(0, rows[1], size=B)    # ~7000 slots per row: 4 chunks each
    ids = torch.from_numpy(np.concatenate([ids0, ids1]).astype(np.int64)).to(DEV)
    go = torch.randn(B, 2 * D, generator=g)
    init = [torch.randn(r, D, generator=g) for r in rows]
    outs = []
    for rep in range(2):
        dw = [w.clone().to(DEV) for w in init]
        s1 = [torch.zeros(r, D, device=DEV) for r in rows]
        tables = ops.TableSet().refresh(dw)
        srt = ops.sort_dedup(tables, lay, ids, None, B)
        args = _lib.OptimArgs(kind=_lib.OPT_ADAGRAD, step=1, lr=0.1, eps=1e-10, beta1=0, beta2=0, weight_decay=0.0, lr_decay=0.0)
        ops.bwd_fused(tables, ops.make_ptr_array(s1), None, lay, B, srt, go.to(DEV), None, args)
        outs.append(([w.cpu() for w in dw], [x.cpu() for x in s1]))
    for a, b in zip(outs[0][0] + outs[0][1], outs[1][0] + outs[1][1]):
        assert torch.equal(a, b), "the chunked reduction must be bit-reproducible"
    for t, idt in enumerate((ids0, ids1)):
        gsum = torch.zeros(rows[t], D, dtype=torch.float64)
        gsum.index_add_(0, torch.from_numpy(idt), go.view(B, 2, D)[:, t].double())
        gabs = torch.zeros(rows[t], D, dtype=torch.float64)
        gabs.index_add_(0, torch.from_numpy(idt), go.view(B, 2, D)[:, t].double().abs())
        got_sum = outs[0][1][t].double().sqrt()   # Adagrad after one step from zero state: sum = g^2
        assert ((got_sum - gsum.abs()).abs() <= 1e-5 * gabs + 1e-6).all(), (t, (got_sum - gsum.abs()).abs().max())
        want_w = init[t].double() - 0.1 * gsum / (gsum.abs() + 1e-10)
        big = gsum.abs() > 1e-3 * gabs            # away from the sign discontinuity of g / |g|
        assert ((outs[0][0][t].double() - want_w).abs()[big] <= 1e-5).all()


def _fused_update_case(opt_name, hp, D, hot):
    B, L = 600, 9
    rows = [50, 3000]
    weights = _tables(rows, D, seed=D + 1)
    specs = [dict(table=0, pooling="mean", mask="pad"), dict(table=1)]
    lay = ops.FeatureLayout([dict(table=0, bag_len=L, pooling="mean", mask="pad"), dict(table=1, bag_len=1)], D, 2)
    dw = [w.clone().to(DEV) for w in weights]
    tables = ops.TableSet().refresh(dw)
    kind = {"sgd": _lib.OPT_SGD, "adagrad": _lib.OPT_ADAGRAD, "rowwise_adagrad": _lib.OPT_ROWWISE_ADAGRAD,
            "lazy_adam": _lib.OPT_LAZY_ADAM}[opt_name]
    if opt_name == "rowwise_adagrad":
        s1 = [torch.zeros(r, device=DEV) for r in rows]
    else:
        s1 = [torch.zeros(r, D, device=DEV) for r in rows]
    s2 = [torch.zeros(r, D, device=DEV) for r in rows]
    p1 = ops.make_ptr_array(s1) if opt_name != "sgd" else None
    p2 = ops.make_ptr_array(s2) if opt_name == "lazy_adam" else None

    ref_w = [w.clone().requires_grad_(True) for w in weights]
    if opt_name == "sgd":
        ropt = torch.optim.SGD(ref_w, lr=hp["lr"])
    elif opt_name == "adagrad":
        ropt = torch.optim.Adagrad(ref_w, lr=hp["lr"], eps=hp["eps"], lr_decay=hp["lr_decay"])
    elif opt_name == "lazy_adam":
        ropt = torch.optim.SparseAdam(ref_w, lr=hp["lr"], betas=(hp["beta1"], hp["beta2"]), eps=hp["eps"])
    else:
        ropt, rstate = None, [torch.zeros(r) for r in rows]

    gsum = [torch.zeros_like(w) for w in weights]
    for step in range(1, 4):
        id_list = [_ids((B, L), rows[0], seed=step, zipf=hot, pad_frac=0.3), _ids((B,), rows[1], seed=10 + step, zipf=hot)]
        ids = torch.cat([i.reshape(-1) for i in id_list]).to(DEV)
        go = torch.randn(B, 2, D, generator=torch.Generator().manual_seed(step))
        out, scale = ops.gather_pool_fwd(tables, lay, ids, None, B, want_scale=True)
        srt = ops.sort_dedup(tables, lay, ids, None, B)
        args = _lib.OptimArgs(kind=kind, step=step, lr=hp["lr"], eps=hp.get("eps", 0.0), beta1=hp.get("beta1", 0.0),
                              beta2=hp.get("beta2", 0.0), weight_decay=0.0, lr_decay=hp.get("lr_decay", 0.0))
        ops.bwd_fused(tables, p1, p2, lay, B, srt, go.view(B, 2 * D).to(DEV), scale, args)
        grads = _dense_grads([w.detach() for w in ref_w], specs, id_list, [None, None], go)
        for t in range(2):
            gsum[t] += grads[t].abs()
        if opt_name == "lazy_adam":
            ropt.zero_grad()
            for w, g in zip(ref_w, grads):
                w.grad = g.to_sparse(1)
            ropt.step()
        elif ropt is not None:
            ropt.zero_grad()
            for w, g in zip(ref_w, grads):
                w.grad = g
            ropt.step()
        else:
            with torch.no_grad():
                for w, s, g in zip(ref_w, rstate, grads):
                    ref_ops.rowwise_adagrad_ref(w, s, g, hp["lr"], hp["eps"])
    for t in range(2):
        got, want = dw[t].cpu(), ref_w[t].detach()
        untouched = gsum[t].sum(1) == 0
        assert torch.equal(got[untouched], weights[t][untouched]), "untouched rows must not move"
        tol = 2e-5 * (want.abs() + 1.0) if opt_name != "sgd" else RTOL * (want.abs() + hp["lr"] * gsum[t])
        assert ((got - want).abs() <= tol + 1e-7).all(), f"table {t}: max err {(got - want).abs().max()}"
    if opt_name == "adagrad":
        for t in range(2):
            np.testing.assert_allclose(s1[t].cpu().numpy(), ropt.state[ref_w[t]]["sum"].numpy(), rtol=2e-5, atol=1e-7)
    if opt_name == "lazy_adam":
        for t in range(2):
            np.testing.assert_allclose(s1[t].cpu().numpy(), ropt.state[ref_w[t]]["exp_avg"].numpy(), rtol=2e-5, atol=2e-6)
            np.testing.assert_allclose(s2[t].cpu().numpy(), ropt.state[ref_w[t]]["exp_avg_sq"].numpy(), rtol=2e-5, atol=1e-9)


def test_segment_sum_equals_dense_backward():
    B, L, D = 300, 5, 16
    rows = [64]
    weights = _tables(rows, D, seed=4)
    id_list = [_ids((B, L), rows[0], seed=8, zipf=True, pad_frac=0.2)]
    specs = [dict(table=0, pooling="sqrtn", mask="pad")]
    lay = ops.FeatureLayout([dict(table=0, bag_len=L, pooling="sqrtn", mask="pad")], D, 1)
    tables = ops.TableSet().refresh([weights[0].to(DEV)])
    ids = id_list[0].reshape(-1).to(DEV)
    go = torch.randn(B, 1, D, generator=torch.Generator().manual_seed(1))
    _, scale = ops.gather_pool_fwd(tables, lay, ids, None, B, want_scale=True)
    srt = ops.sort_dedup(tables, lay, ids, None, B)
    rg = ops.segment_sum(lay, B, srt, go.view(B, D).to(DEV), scale).cpu()
    n = int(srt.n_seg.item())
    keys = _u32(srt.sorted_keys.cpu())[srt.seg_start.cpu()[:n].long()]
    dense = _dense_grads(weights, specs, id_list, [None], go)[0]
    got = torch.zeros_like(dense)
    valid = keys != 0xFFFFFFFF
    got[keys[valid]] = rg[:n][valid]
    np.testing.assert_allclose(got.numpy(), dense.numpy(), rtol=1e-4, atol=1e-5)
    # run-to-run bit reproducibility (fixed reduction order)
    rg2 = ops.segment_sum(lay, B, ops.sort_dedup(tables, lay, ids, None, B), go.view(B, D).to(DEV), scale).cpu()
    assert torch.equal(rg[:n], rg2[:n])


# ------------------------------------------------------------------------------------------ FM second order
@pytest.mark.parametrize("B,F,D", [(1, 2, 4), (257, 26, 16), (64, 39, 32), (33, 5, 64), (100, 40, 128), (50, 7, 10), (31, 3, 1)])
def test_fm2_forward_backward(B, F, D):
    g = torch.Generator().manual_seed(B + F + D)
    v = torch.randn(B, F, D, generator=g)
    gy = torch.randn(B, generator=g)
    vr = v.clone().requires_grad_(True)
    yr = ref_ops.fm2_ref(vr)
    yr.backward(gy)
    vd = v.to(DEV).requires_grad_(True)
    y = ops.fm2(vd)
    y.backward(gy.to(DEV))
    scale = 0.5 * ((v.sum(1) ** 2).sum(-1) + (v * v).sum((1, 2)))  # sum of |terms|
    assert ((y.detach().cpu() - yr.detach()).abs() <= RTOL * scale + 1e-6).all()
    gscale = gy.abs().view(B, 1, 1) * (v.abs().sum(1, keepdim=True) + v.abs())
    assert ((vd.grad.cpu() - vr.grad).abs() <= RTOL * gscale + 1e-6).all()


def test_fm2_backward_fused_accumulate():
    B, F, D = 130, 26, 16
    g = torch.Generator().manual_seed(0)
    v, gy, gin = torch.randn(B, F, D, generator=g), torch.randn(B, generator=g), torch.randn(B, F, D, generator=g)
    base = ops.fm2_bwd(v.to(DEV), gy.to(DEV))
    fused = ops.fm2_bwd(v.to(DEV), gy.to(DEV), gin.to(DEV))
    np.testing.assert_allclose(fused.cpu().numpy(), (base.cpu() + gin).numpy(), rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------------------------------ full-size properties
def test_full_size_properties_deepfm_config():
    """cfg2 shape (26 tables x 1e6 rows x D16, B 16384): properties that need no CPU-sized oracle."""
    F, R, D, B = 26, 1_000_000, 16, 16384
    gen = torch.Generator(device=DEV).manual_seed(2020)
    weights = [torch.randn(R, D, device=DEV, generator=gen) for _ in range(F)]
    ids2d = torch.randint(0, R, (F, B), device=DEV, generator=gen)
    lay = ops.FeatureLayout([dict(table=f, bag_len=1) for f in range(F)], D, F)
    tables = ops.TableSet().refresh(weights)
    ids = ids2d.reshape(-1).contiguous()
    out, _ = ops.gather_pool_fwd(tables, lay, ids, None, B)
    # gather == indexing (bit-exact), checked per table on the device
    out3 = out.view(B, F, D)
    for f in (0, 7, 25):
        assert torch.equal(out3[:, f], weights[f][ids2d[f]])
    srt = ops.sort_dedup(tables, lay, ids, None, B)
    N = srt.N
    keys = _u32(srt.sorted_keys)[:N].view(F, B)
    assert (keys[:, 1:] >= keys[:, :-1]).all(), "each table's keys must be sorted"
    perm = srt.perm[:N].long()
    assert torch.equal(torch.sort(perm).values, torch.arange(N, device=DEV)), "perm must be a permutation"
    assert torch.equal(ids[perm], _u32(srt.sorted_keys)[:N])
    n_seg = int(srt.n_seg.item())
    assert n_seg == sum(int(torch.unique(ids2d[f]).numel()) for f in range(F))
    # SGD linearity: sum of all weights moves by -lr * sum of all upstream gradients
    go = torch.randn(B, F * D, device=DEV, generator=gen)
    before = torch.stack([w.double().sum() for w in weights]).sum()
    args = _lib.OptimArgs(kind=_lib.OPT_SGD, step=1, lr=0.5, eps=0, beta1=0, beta2=0, weight_decay=0, lr_decay=0)
    ops.bwd_fused(tables, None, None, lay, B, srt, go, None, args)
    after = torch.stack([w.double().sum() for w in weights]).sum()
    np.testing.assert_allclose((before - after).item(), 0.5 * go.double().sum().item(), rtol=1e-3, atol=2.0)


# ------------------------------------------------------------------------------------------ DCN cross (tcgen05)
def _bf16(t):
    return t.to(torch.bfloat16)


@pytest.fixture(params=["pair", "single"])
def dcn_kernel(request):
    """K5 runs on the CTA-pair GEMM (default) or on the single-CTA kernel of round 1; the tests cover both."""
    ops.set_dcn_2sm(request.param == "pair")
    yield request.param
    ops.set_dcn_2sm(True)


@pytest.mark.parametrize("B,d", [(128, 128), (256, 64), (1000, 848), (4096, 896), (130, 72), (33000, 848), (8, 8),
                                 (777, 1032)])
def test_dcn_cross_fwd_matches_fp32_reference(B, d, dcn_kernel):
    """bf16 operands, fp32 accumulate: compare with the same bf16-rounded inputs multiplied in fp32 (tolerance 1e-2)."""
    g = torch.Generator().manual_seed(B + d)
    x0, xl = torch.randn(B, d, generator=g) * 0.5, torch.randn(B, d, generator=g) * 0.5
    W, b = torch.randn(d, d, generator=g) / d ** 0.5, torch.randn(d, generator=g) * 0.1
    x0b, xlb, Wb = _bf16(x0), _bf16(xl), _bf16(W)
    out, u = ops.dcn_cross_fwd(xlb.to(DEV), x0b.to(DEV), Wb.to(DEV), b.to(DEV))
    u_ref = xlb.float() @ Wb.float().t() + b
    out_ref = x0b.float() * u_ref + xlb.float()
    np.testing.assert_allclose(u.float().cpu().numpy(), u_ref.numpy(), rtol=1e-2, atol=1e-2)
    np.testing.assert_allclose(out.float().cpu().numpy(), out_ref.numpy(), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("B,d", [(256, 128), (1000, 848), (33000, 848), (130, 72), (777, 1032)])
def test_dcn_cross_dgrad_wgrad(B, d, dcn_kernel):
    g = torch.Generator().manual_seed(B * 3 + d)
    x0, xl, go = (torch.randn(B, d, generator=g) * 0.5 for _ in range(3))
    W = torch.randn(d, d, generator=g) / d ** 0.5
    x0b, xlb, gob, Wb = _bf16(x0), _bf16(xl), _bf16(go), _bf16(W)
    gub = _bf16(gob.float() * x0b.float())
    gx, prev = ops.dcn_cross_dgrad(gub.to(DEV), Wb.t().contiguous().to(DEV), gob.to(DEV), x0b.to(DEV))
    gx_ref = gub.float() @ Wb.float() + gob.float()
    np.testing.assert_allclose(gx.float().cpu().numpy(), gx_ref.numpy(), rtol=1e-2, atol=1e-2)
    np.testing.assert_allclose(prev.float().cpu().numpy(), (gx_ref * x0b.float()).numpy(), rtol=2e-2, atol=1e-2)
    gw = ops.dcn_cross_wgrad(gub.to(DEV), xlb.to(DEV))
    gw_ref = gub.float().t() @ xlb.float()
    np.testing.assert_allclose(gw.cpu().numpy(), gw_ref.numpy(), rtol=1e-2, atol=1e-2 * B ** 0.5 * 0.25)


def test_cross_net_autograd_matches_oracle(dcn_kernel):
    from oracle import ref_models
    B, d, L = 512, 845, 3
    g = torch.Generator().manual_seed(1)
    x0 = torch.randn(B, d, generator=g) * 0.3
    Ws = [torch.randn(d, d, generator=g) / d ** 0.5 for _ in range(L)]
    bs = [torch.randn(d, generator=g) * 0.1 for _ in range(L)]
    go = torch.randn(B, d, generator=g)
    x0r = x0.clone().requires_grad_(True)
    Wr, br = [w.clone().requires_grad_(True) for w in Ws], [b.clone().requires_grad_(True) for b in bs]
    yr = ref_models.cross_net_ref(x0r, Wr, br)
    yr.backward(go)
    x0d = x0.to(DEV).requires_grad_(True)
    Wd, bd = [w.to(DEV).requires_grad_(True) for w in Ws], [b.to(DEV).requires_grad_(True) for b in bs]
    y = ops.cross_net(x0d, Wd, bd)
    y.backward(go.to(DEV))

    def close(a, b, what):
        a, b = a.detach().cpu().float(), b.detach().float()
        err = (a - b).norm() / b.norm()
        assert err < 1e-2, f"{what}: relative error {err:.4f}"

    close(y, yr, "out")
    close(x0d.grad, x0r.grad, "grad x0")
    for l in range(L):
        close(Wd[l].grad, Wr[l].grad, f"grad W{l}")
        close(bd[l].grad, br[l].grad, f"grad b{l}")


@pytest.mark.parametrize("B,d", [(64, 8), (1000, 845), (4099, 77), (513, 1030)])
def test_cross_net_head_matches_unfused_head(B, d):
    """cross_net_head(x0, W, b, w) == cross_net(x0, W, b) @ w: same bf16 chain, the row dot and its backward fused
    (the head reads / writes the chain's bf16 tensors, so forward and gradients agree to fp32 summation order, and the
    gradients that pass through bf16(g_y (x) w) agree exactly in their bf16 operands)."""
    g = torch.Generator().manual_seed(B + 7 * d)
    L = 2
    x0 = (torch.randn(B, d, generator=g) * 0.3).to(DEV)
    Ws = [(torch.randn(d, d, generator=g) / d ** 0.5).to(DEV) for _ in range(L)]
    bs = [(torch.randn(d, generator=g) * 0.1).to(DEV) for _ in range(L)]
    w = (torch.randn(d, generator=g) / d ** 0.5).to(DEV)
    gy = torch.randn(B, generator=g).to(DEV)

    def run(fused):
        leaves = [t.clone().requires_grad_(True) for t in (x0, w, *Ws, *bs)]
        x, hw, W, b = leaves[0], leaves[1], leaves[2:2 + L], leaves[2 + L:]
        y = ops.cross_net_head(x, W, b, hw) if fused else ops.cross_net(x, W, b) @ hw
        y.backward(gy)
        return y.detach(), [t.grad for t in leaves]

    y1, g1 = run(True)
    y0, g0 = run(False)
    scale = y0.abs().max().item()
    assert (y1 - y0).abs().max().item() <= 1e-5 * max(scale, 1.0) * d ** 0.5
    for a, b_, what in zip(g1, g0, ["x0", "head w"] + [f"W{l}" for l in range(L)] + [f"b{l}" for l in range(L)]):
        err = ((a - b_).norm() / b_.norm().clamp_min(1e-30)).item()
        assert err < 2e-5, f"grad {what}: relative error {err:.2e}"


def test_dcn_head_kernels_match_fp32_reference():
    lib = _lib.load()
    B, d = 777, 845
    dp = (d + 7) // 8 * 8
    g = torch.Generator().manual_seed(5)
    xl = torch.zeros(B, dp, dtype=torch.bfloat16)
    x0 = torch.zeros(B, dp, dtype=torch.bfloat16)
    xl[:, :d] = (torch.randn(B, d, generator=g) * 0.5).to(torch.bfloat16)
    x0[:, :d] = (torch.randn(B, d, generator=g) * 0.5).to(torch.bfloat16)
    w, gy = torch.randn(d, generator=g), torch.randn(B, generator=g)
    xld, x0d, wd, gyd = xl.to(DEV), x0.to(DEV), w.to(DEV), gy.to(DEV)
    y = torch.empty(B, device=DEV)
    _lib.check(lib.ptrec_dcn_head_fwd(xld.data_ptr(), B, d, dp, wd.data_ptr(), y.data_ptr(), None), "head_fwd")
    y_ref = xl[:, :d].double() @ w.double()
    np.testing.assert_allclose(y.cpu().double().numpy(), y_ref.numpy(), rtol=1e-5, atol=1e-5 * d ** 0.5)
    g_out = torch.full((B, dp), 7.0, dtype=torch.bfloat16, device=DEV)
    g_u = torch.full((B, dp), 7.0, dtype=torch.bfloat16, device=DEV)
    gw = torch.empty(d, device=DEV)
    nb = lib.ptrec_dcn_bwd_layer_workspace_bytes(B, dp)
    ws = torch.empty(nb, dtype=torch.uint8, device=DEV)
    _lib.check(lib.ptrec_dcn_head_bwd(gyd.data_ptr(), wd.data_ptr(), xld.data_ptr(), x0d.data_ptr(), B, d, dp,
                                      g_out.data_ptr(), g_u.data_ptr(), gw.data_ptr(), ws.data_ptr(), nb, None), "head_bwd")
    go_ref = torch.zeros(B, dp, dtype=torch.bfloat16)
    go_ref[:, :d] = (gy[:, None] * w[None, :]).to(torch.bfloat16)   # fp32 product rounded to bf16: bit-exact
    assert torch.equal(g_out.cpu(), go_ref), "g_out = bf16(g_y (x) w), zero padded"
    gu_ref = (go_ref.float() * x0.float()).to(torch.bfloat16)
    assert torch.equal(g_u.cpu(), gu_ref), "g_u = bf16(g_out * x0)"
    gw_ref = (gy.double()[:, None] * xl[:, :d].double()).sum(0)
    np.testing.assert_allclose(gw.cpu().double().numpy(), gw_ref.numpy(), rtol=1e-5, atol=1e-5 * B ** 0.5)


@pytest.mark.parametrize("B,d", [(50, 13), (1000, 845), (33, 1030)])
def test_dcn_pack_unpack_final_unaligned_rows(B, d):
    """The fp32 side of the glue kernels has rows of d floats (any d): pack / unpack round trip and bwd_final are exact."""
    lib = _lib.load()
    dp = (d + 7) // 8 * 8
    g = torch.Generator().manual_seed(d)
    x = torch.randn(B, d, generator=g).to(DEV)
    xb = torch.full((B, dp), 3.0, dtype=torch.bfloat16, device=DEV)
    _lib.check(lib.ptrec_dcn_pack_input(x.data_ptr(), d, B, d, dp, xb.data_ptr(), None), "pack")
    assert torch.equal(xb[:, :d], x.to(torch.bfloat16)) and (xb[:, d:] == 0).all()
    out = torch.empty(B, d, device=DEV)
    _lib.check(lib.ptrec_dcn_unpack(xb.data_ptr(), B, d, dp, out.data_ptr(), None), "unpack")
    assert torch.equal(out, x.to(torch.bfloat16).float())
    gx0 = torch.randn(B, dp, generator=g).to(DEV)
    fin = torch.empty(B, d, device=DEV)
    _lib.check(lib.ptrec_dcn_bwd_final(gx0.data_ptr(), xb.data_ptr(), B, d, dp, fin.data_ptr(), None), "final")
    assert torch.equal(fin, gx0[:, :d] + xb[:, :d].float())
    go = torch.empty(B, dp, dtype=torch.bfloat16, device=DEV)
    gu = torch.empty(B, dp, dtype=torch.bfloat16, device=DEV)
    _lib.check(lib.ptrec_dcn_bwd_init(x.data_ptr(), d, xb.data_ptr(), B, d, dp, go.data_ptr(), gu.data_ptr(), None), "init")
    assert torch.equal(go, xb) and torch.equal(gu, (xb.float() * xb.float()).to(torch.bfloat16))


# ------------------------------------------------------------------------------------------ DIN attention pooling
@pytest.fixture(params=["simt", "tensor_core_fwd", "tensor_core_fwd_bwd"])
def din_build(request):
    """K4 has fp32 CUDA-core kernels and tcgen05 kernels (fp16 x 2 operand planes) for forward and backward: same
    contract."""
    lib = _lib.load()
    before = lib.ptrec_din_tc_enabled()
    lib.ptrec_set_din_tc({"simt": 0, "tensor_core_fwd": 1, "tensor_core_fwd_bwd": 3}[request.param])
    yield request.param
    lib.ptrec_set_din_tc(before)


@pytest.mark.parametrize("B,L,DQ,hidden", [(3, 5, 32, (80, 40)), (64, 100, 32, (80, 40)), (33, 130, 16, (80, 40)),
                                            (40, 50, 32, (64, 32)), (17, 100, 16, (64, 32)), (300, 300, 32, (80, 40))])
def test_din_attention_pool_forward_backward(B, L, DQ, hidden, din_build):
    from oracle import ref_models
    g = torch.Generator().manual_seed(B + L + DQ)
    H1, H2 = hidden
    fc1, fc2, fc3 = torch.nn.Linear(4 * DQ, H1), torch.nn.Linear(H1, H2), torch.nn.Linear(H2, 1)
    with torch.no_grad():
        for m in (fc1, fc2, fc3):
            m.weight.copy_(torch.randn(m.weight.shape, generator=g) * 0.2)
            m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
    q = torch.randn(B, DQ, generator=g)
    keys = torch.randn(B, L, DQ, generator=g)
    lens = torch.randint(0, L + 1, (B,), generator=g)
    lens[0] = L
    go = torch.randn(B, DQ, generator=g)
    qr, kr = q.clone().requires_grad_(True), keys.clone().requires_grad_(True)
    out_ref = ref_models.din_attention_ref(qr, kr, lens, fc1, fc2, fc3)
    out_ref.backward(go)
    params = [p.detach().clone().to(DEV).requires_grad_(True) for p in (fc1.weight, fc1.bias, fc2.weight, fc2.bias, fc3.weight, fc3.bias)]
    qd, kd = q.to(DEV).requires_grad_(True), keys.to(DEV).requires_grad_(True)
    out = ops.din_attn_pool(qd, kd, lens.to(DEV), *params)
    out.backward(go.to(DEV))

    def close(a, b, what, rtol=1e-5):
        a, b = a.detach().cpu(), b.detach()
        tol = rtol * b.abs().max().clamp(min=1e-3) * 20 + 1e-6  # sums of ~L*H terms: scale by the tensor's magnitude
        assert ((a - b).abs() <= tol).all(), f"{what}: max err {(a - b).abs().max():.3e} tol {tol:.3e}"

    close(out, out_ref, "pooled")
    close(qd.grad, qr.grad, "grad q")
    close(kd.grad, kr.grad, "grad keys")
    for p, r, n in zip(params, (fc1.weight, fc1.bias, fc2.weight, fc2.bias, fc3.weight, fc3.bias),
                       ("W1", "b1", "W2", "b2", "W3", "b3")):
        close(p.grad, r.grad, "grad " + n)
    # strided q / keys views of one [B, 1+L, DQ] buffer (the layout the DIN model uses)
    seq = torch.cat([q.unsqueeze(1), keys], dim=1).to(DEV)
    out2, scores = ops.din_attn_pool_fwd(seq[:, 0], seq[:, 1:], lens.to(DEV), [p.detach() for p in params], want_scores=True)
    assert torch.equal(out2, out.detach())
    assert (scores.cpu()[torch.arange(L).unsqueeze(0) >= lens.unsqueeze(1)] == 0).all()


# ---------------------------------------------------------------------------------------------------------------
# K6 fp32-faithful Linear on tcgen05 (bf16 x 3 split operands)
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(params=["2sm", "1sm"])
def tc_mode(request):
    """K6 GEMM: CTA-pair (cta_group::2) and single-CTA kernels must agree with fp64 alike."""
    lib = _lib.load()
    lib.ptrec_tc_set_2sm(1 if request.param == "2sm" else 0)
    yield request.param
    lib.ptrec_tc_set_2sm(1)


def _planes_sum(pl, rows, cols):
    return (pl[2, :rows, :cols].float() + pl[1, :rows, :cols].float()) + pl[0, :rows, :cols].float()


@pytest.mark.gpu
@pytest.mark.parametrize("R,C", [(64, 64), (200, 429), (1031, 13), (4096, 400)])
def test_tc_split3_planes_are_an_exact_decomposition(R, C):
    gen = torch.Generator().manual_seed(R + C)
    x = (torch.randn(R, C, generator=gen) * torch.exp(3 * torch.randn(R, C, generator=gen))).to(DEV)
    y = torch.randn(R, C, generator=gen).to(DEV)
    pl, plt, cs = ops.tc_split3(x, want_planes=True, want_t=True, want_colsum=True)
    assert torch.equal(_planes_sum(pl, R, C), x)                      # x0 + x1 + x2 == x bit for bit
    assert torch.equal(_planes_sum(plt, C, R), x.t())
    assert (pl[:, :, C:] == 0).all() and (plt[:, :, R:] == 0).all()   # pad columns are zero
    ref = x.double().sum(0)
    assert torch.allclose(cs.double(), ref, rtol=1e-5, atol=1e-5 * x.abs().sum(0).max().item())
    # fused ReLU backward: g * (y > 0)
    pl2, _, cs2 = ops.tc_split3(x, relu_ref=y, want_colsum=True)
    masked = x * (y > 0)
    assert torch.equal(_planes_sum(pl2, R, C), masked)
    assert torch.allclose(cs2.double(), masked.double().sum(0), rtol=1e-5, atol=1e-5 * x.abs().sum(0).max().item())
    # strided source (a column slice)
    wide = torch.randn(R, C + 5, generator=gen).to(DEV)
    pl3, _, _ = ops.tc_split3(wide[:, 3:3 + C])
    assert torch.equal(_planes_sum(pl3, R, C), wide[:, 3:3 + C])


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 400, 429), (1000, 16, 40), (2048, 429, 400), (130, 1, 700)])
def test_tc_gemm_split3_has_fp32_level_error(M, N, K, tc_mode):
    """|C - C_fp64| is at the level of an fp32 GEMM (<= 2e-6 of sum|a||b|), far from bf16 / tf32 error."""
    gen = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=gen).to(DEV)
    b = torch.randn(N, K, generator=gen).to(DEV)
    bias = torch.randn(N, generator=gen).to(DEV)
    pa, _, _ = ops.tc_split3(a)
    pb, _, _ = ops.tc_split3(b)
    ref = a.double() @ b.double().t()
    scale = (a.double().abs() @ b.double().abs().t())
    out = ops.tc_gemm_split3(pa, pb, K)
    assert out.shape == (M, N)
    err = ((out.double() - ref).abs() / scale).max().item()
    fp32_err = ((a @ b.t()).double() - ref).abs().div(scale).max().item()
    assert err <= max(2e-7, 2 * fp32_err), (err, fp32_err)
    out2 = ops.tc_gemm_split3(pa, pb, K, bias=bias, relu=True)
    ref2 = torch.relu(ref + bias.double())
    assert ((out2.double() - ref2).abs() / (scale + bias.abs().double())).max().item() <= max(2e-7, 2 * fp32_err)
    # split-K (the weight-gradient shape): same answer up to summation order
    out3 = ops.tc_gemm_split3(pa, pb, K, splits=3)
    assert ((out3.double() - ref).abs() / scale).max().item() <= max(3e-7, 2 * fp32_err)


@pytest.mark.gpu
def test_tc_gemm_split3_weight_gradient_shape_with_transposed_planes(tc_mode):
    B, N, K = 4096, 400, 429
    gen = torch.Generator().manual_seed(5)
    g = torch.randn(B, N, generator=gen).to(DEV)
    x = torch.randn(B, K, generator=gen).to(DEV)
    _, gt, _ = ops.tc_split3(g, want_planes=False, want_t=True)
    _, xt, _ = ops.tc_split3(x, want_planes=False, want_t=True)
    dw = ops.tc_gemm_split3(gt, xt, B, splits=0)
    ref = g.double().t() @ x.double()
    scale = g.double().abs().t() @ x.double().abs()
    assert dw.shape == (N, K)
    assert ((dw.double() - ref).abs() / scale).max().item() <= 3e-7


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,K", [(4096, 400, 429), (1000, 128, 64), (777, 16, 1030), (64, 200, 13)])
def test_tc_gemm_split3_tn_weight_gradient_from_row_major_planes(B, N, K, tc_mode):
    """dW = g^T x straight from the row-major planes (MN-major tcgen05 operands): no transposed copies."""
    gen = torch.Generator().manual_seed(B + N)
    g = torch.randn(B, N, generator=gen).to(DEV)
    x = torch.randn(B, K, generator=gen).to(DEV)
    pg, _, _ = ops.tc_split3(g)
    px, _, _ = ops.tc_split3(x)
    ref = g.double().t() @ x.double()
    scale = g.double().abs().t() @ x.double().abs()
    fp32_err = ((g.t() @ x).double() - ref).abs().div(scale).max().item()
    for splits, tol in ((0, max(3e-7, 2 * fp32_err)), (1, 1e-6)):  # one split = one long chain of truncating adds
        dw = ops.tc_gemm_split3_tn(pg, N, px, K, splits=splits)
        assert dw.shape == (N, K)
        assert ((dw.double() - ref).abs() / scale).max().item() <= tol, (splits, fp32_err)


# ---------------------------------------------------------------------------------------------------------------
# K8 FM head + row dot
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("B,F,D,nd", [(257, 26, 16, 13), (64, 3, 4, 0), (1000, 26, 64, 13), (33, 5, 8, 1)])
def test_fm_head_forward_backward_match_torch(B, F, D, nd):
    from pytorchrec_b200.model.layer.interaction import fm_head
    gen = torch.Generator().manual_seed(B + F)
    v = torch.randn(B, F, D, generator=gen).to(DEV).requires_grad_(True)
    w1 = torch.randn(B, F, 1, generator=gen).to(DEV).requires_grad_(True)
    x = torch.randn(B, nd, generator=gen).to(DEV) if nd else None
    wd = torch.randn(1, nd, generator=gen).to(DEV).requires_grad_(True) if nd else None
    bias = torch.randn((), generator=gen).to(DEV).requires_grad_(True)
    gl = torch.randn(B, generator=gen).to(DEV)
    gd = torch.randn(B, F * D + nd, generator=gen).to(DEV)

    def ref():
        s = v.sum(1)
        logit = w1.sum(dim=(1, 2)) + 0.5 * (s * s - (v * v).sum(1)).sum(1) + bias
        if nd:
            logit = logit + (x @ wd.t()).squeeze(-1)
        flat = v.reshape(B, -1)
        return logit, (torch.cat([flat, x], 1) if nd else flat)

    outs = []
    for fn in (ref, lambda: fm_head(v, w1, x, wd, bias, True)):
        logit, deep_in = fn()
        torch.autograd.backward([logit, deep_in], [gl, gd])
        outs.append([logit.detach(), deep_in.detach(), v.grad, w1.grad, bias.grad] + ([wd.grad] if nd else []))
        v.grad = w1.grad = bias.grad = None
        if nd:
            wd.grad = None
    names = ["logit", "deep_in", "gv", "gw1", "gbias", "gwd"]
    for n, a, b in zip(names, outs[0], outs[1]):
        assert a.shape == b.shape, n
        scale = max(1.0, a.abs().max().item())
        tol = (2e-5 if n in ("gbias", "gwd") else 3e-6) * scale * (D if n == "logit" else 1)
        assert (a - b).abs().max().item() <= tol, (n, (a - b).abs().max().item(), tol)
    assert torch.equal(outs[0][1], outs[1][1])  # the tower input is a copy: bit-exact


@pytest.mark.gpu
@pytest.mark.parametrize("B,H", [(1000, 400), (64, 1024), (257, 32), (3000, 1024), (500, 640)])
def test_row_dot_matches_linear(B, H):
    from pytorchrec_b200.model.layer.interaction import row_dot
    gen = torch.Generator().manual_seed(H)
    h = torch.randn(B, H, generator=gen).to(DEV).requires_grad_(True)
    w = torch.randn(1, H, generator=gen).to(DEV).requires_grad_(True)
    g = torch.randn(B, generator=gen).to(DEV)
    res = []
    for fn in (lambda: torch.nn.functional.linear(h, w).squeeze(-1), lambda: row_dot(h, w)):
        y = fn()
        y.backward(g)
        res.append((y.detach(), h.grad, w.grad))
        h.grad = w.grad = None
    for a, b in zip(*res):
        assert a.shape == b.shape
        assert (a - b).abs().max().item() <= 2e-5 * max(1.0, a.abs().max().item())


@pytest.mark.gpu
def test_producer_written_planes_are_the_exact_split_of_the_fp32_result(tc_mode):
    """GEMM epilogue and FM head write the next layer's bf16 planes themselves: same bits as a split pass."""
    gen = torch.Generator().manual_seed(9)
    a = torch.randn(700, 429, generator=gen).to(DEV)
    w = torch.randn(400, 429, generator=gen).to(DEV)
    b = torch.randn(400, generator=gen).to(DEV)
    pa, pw = ops.tc_split3(a)[0], ops.tc_split3(w)[0]
    y, py = ops.tc_gemm_split3(pa, pw, 429, bias=b, relu=True, want_planes=True)
    assert torch.equal(y, ops.tc_gemm_split3(pa, pw, 429, bias=b, relu=True))
    assert torch.equal(py, ops.tc_split3(y)[0])
    w2 = torch.randn(52, 429, generator=gen).to(DEV)      # N = 52: pitch 52 (fp32) vs 56 (planes), pad must be 0
    y2, py2 = ops.tc_gemm_split3(pa, ops.tc_split3(w2)[0], 429, want_planes=True)
    assert torch.equal(py2, ops.tc_split3(y2)[0])
    v = torch.randn(300, 26 * 16, generator=gen).to(DEV)
    x = torch.randn(300, 13, generator=gen).to(DEV)
    logit, deep_in, planes = ops.fm_head_fwd(v, None, x, None, None, 26, 16, True, True)
    assert torch.equal(deep_in, torch.cat([v, x], 1)) and torch.equal(planes, ops.tc_split3(deep_in)[0])


# ------------------------------------------------------------------------------------------ K9 loss
@pytest.mark.parametrize("n", [1, 5, 4096, 16384, 300001])
def test_bce_logits_mean_matches_torch(n):
    """K9 against ``torch.nn.BCEWithLogitsLoss()`` (the loss IModel.train_step applies, torchrec/model/IModel.py:121):
    value and gradient, saturated logits included; bit-reproducible from run to run."""
    gen = torch.Generator().manual_seed(n)
    x = (4.0 * torch.randn(n, generator=gen))
    x[::7] = 60.0 * torch.sign(x[::7])          # saturated both ways
    t = (torch.rand(n, generator=gen) < 0.3).float()
    xd = x.to(DEV).requires_grad_(True)
    loss = ops.bce_logits_mean(xd, t.to(DEV))
    (2.5 * loss).backward()
    xr = x.double().requires_grad_(True)
    ref = torch.nn.BCEWithLogitsLoss()(xr, t.double())
    (2.5 * ref).backward()
    assert abs(loss.item() - ref.item()) <= 1e-6 * max(1.0, abs(ref.item()))
    np.testing.assert_allclose(xd.grad.cpu().double().numpy(), xr.grad.numpy(), rtol=2e-6, atol=1e-7 / n)
    again = ops.bce_logits_mean(x.to(DEV), t.to(DEV))
    assert again.item() == loss.item()
    if n > 2:
        assert ops.bce_logits_mean(x.to(DEV)[::2], t.to(DEV)[::2]) is None   # non-contiguous: the caller uses the module


@pytest.mark.parametrize("B,L,hidden,interleaved", [(37, 100, (80, 40), False), (300, 50, (64, 32), True), (5, 130, (80, 40), True)])
def test_din_keys_by_id_equal_gathered_keys(B, L, hidden, interleaved):
    """K4 with the gather fused in: reading key (b, l) from the two tables by id gives bit-identical outputs and gradients
    to the same kernels fed with the gathered [B, L, DQ] tensor (same arithmetic, only the address of the rows differs);
    tables interleaved with optimizer state (row pitch 2D) included; an id out of range raises the error word."""
    lib = _lib.load()
    if lib.ptrec_din_tc_enabled() != 3:
        pytest.skip("tensor-core K4 builds switched off")
    D, DQ = 16, 32
    H1, H2 = hidden
    g = torch.Generator().manual_seed(B * L)
    rows = (1000, 57)
    store = [torch.randn(r, 2 * D if interleaved else D, generator=g).to(DEV) for r in rows]
    tabs = [st[:, :D] for st in store]
    ids = [torch.randint(0, r, (B, 1 + L), generator=g).to(DEV) for r in rows]
    lens = torch.randint(0, L + 1, (B,), generator=g).to(DEV)
    lens[0] = L
    params = [(torch.randn(sh, generator=g) * 0.2).to(DEV) for sh in ((H1, 4 * DQ), (H1,), (H2, H1), (H2,), (1, H2), (1,))]
    seq = torch.cat([tabs[0][ids[0]], tabs[1][ids[1]]], dim=2)           # [B, 1 + L, DQ]
    q, keys = seq[:, 0].contiguous(), seq[:, 1:].contiguous()
    gp = torch.randn(B, DQ, generator=g).to(DEV)
    out_ref, _ = ops.din_attn_pool_fwd(q, keys, lens, params)
    gq_ref, gk_ref, gpar_ref = ops.din_attn_pool_bwd(q, keys, lens, params, gp)
    flat_ids = [i.reshape(-1).contiguous() for i in ids]
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    out = ops.din_attn_pool_fwd_ids(q, tabs, flat_ids, 1 + L, 1, err, lens, L, params)
    assert torch.equal(out, out_ref) and int(err.item()) == 0
    g_seq = torch.full((B, 1 + L, DQ), 7.0, device=DEV)
    gq, gpar = ops.din_attn_pool_bwd_ids(q, tabs, flat_ids, 1 + L, 1, lens, L, params, gp, g_seq[:, 1:])
    assert torch.equal(gq, gq_ref) and torch.equal(g_seq[:, 1:], gk_ref)
    assert (g_seq[:, 0] == 7.0).all(), "row 0 of the buffer belongs to the caller"
    for a, b in zip(gpar, gpar_ref):
        assert torch.equal(a, b)
    bad = [i.clone() for i in flat_ids]
    bad[1][1] = rows[1]                                                   # sample 0, position 0 (< lens[0])
    ops.din_attn_pool_fwd_ids(q, tabs, bad, 1 + L, 1, err, lens, L, params)
    assert int(err.item()) == 1
