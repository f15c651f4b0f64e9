"""CPU oracle for the ms+cs dense contrastive loss -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker / timed CPU baseline.  The product path (``mscs_b200``) never imports it and fails
loudly when its CUDA library is missing.

Parity status: PINNED against the executable reference.  The reference ships no tests or
golden vectors for this path, so the oracle is pinned against outputs of the
reference itself: ``tests/golden/make_golden.py`` imports ``/root/reference/losses/*`` in the
build container (stubbed ``utils`` package, ``Tensor.cuda`` neutralised), runs it on the
synthetic inputs of ``mscs_b200.synth`` and commits the results under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every oracle function against those fixtures.

Modules
  mt19937.py     MT19937 + the Fisher-Yates ``torch.randperm`` CPU algorithm (numpy)
  sampling.py    nearest label down-sampling, class histograms, pair list, V rule, indices (numpy)
  loss_fp64.py   chunked fp64 loss + analytic gradients (numpy)
  torch_port.py  fp32 torch restatement with autograd -- the timed CPU baseline ("port")
"""
