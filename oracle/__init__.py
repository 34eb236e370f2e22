"""CPU oracle for the pytorchrec_b200 hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pytorchrec_b200/`` imports this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may.
It restates, in plain torch / numpy on the CPU, what the reference (Troublem1/PyTorchRec, mounted
read-only at /root/reference when golden vectors are generated) computes on this path; each
function cites the reference file:line it follows.

Pinning status
  * PINNED by executing the unmodified reference here (``oracle/make_golden.py``, ``make_golden_ncf.py``,
    ``make_golden_reader.py`` -> ``tests/golden/*.npz``): the NCF model with its MLP / Dense tower, the data readers'
    batch assembly and pair-wise negative sampler (``ref_reader.py``), and
    the lifecycle / init / param groups / train_step restatement (``IModelRef``), ``nn.Embedding`` gather,
    the SVD++ masked ``sum / sqrt(count)`` pooling, the SASRec masked-mean idiom, the FunkSVD / SVD++ /
    NCF-style interactions, dense SGD / Adam / AdamW steps, ``CrossedColumn`` arithmetic.
  * PARITY UNPINNED by the reference (it has neither tests nor these components):
    FM / DeepFM / DCN-v2 / DIN models, Adagrad / row-wise Adagrad / lazy Adam updates, sort/dedup
    artefacts.  Their oracle is a restatement in the reference's idiom whose arithmetic lives in torch
    2.11 CPU (``torch.sort(stable=True)``, ``torch.unique``, ``torch.optim.Adagrad``, ``torch.optim.SparseAdam``);
    golden files record the torch version.
"""
