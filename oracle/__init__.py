"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's hot path.

Nothing under oracle/ is part of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` leg may import it, and only as the checker or the timed CPU baseline.  The product path
(hp-vae-gan_b200/) never imports it and has no CPU fallback.

  oracle/port.py    functional PyTorch-CPU (fp32/fp64) restatement of modules/networks_3d.py, networks_2d.py,
                    losses.py, utils.py and the utils/images.py functions they call, driven by a plain state_dict;
                    storage('bf16') adds the precision model of the CUDA path (bf16-stored activations/gradients).
  oracle/train_ref.py  the per-scale training loop (train_video.py:44-88, 111-202) on top of port.py.
  oracle/np_ops.py  numpy restatement of the primitive operators the path resolves to inside PyTorch (the
                    reference's arithmetic lives in its third-party dependency torch, pinned at pytorch==1.4.0 in
                    env.sh:3; installed here: torch 2.11.0): convolution, training-mode batch norm, leaky ReLU,
                    align_corners linear resize, spectral-norm power iteration.

Parity pin: the reference has no tests or golden vectors of its own, so tests/golden/*.pt were produced by
tests/golden/make_golden.py, which imports the UNMODIFIED reference modules from /root/reference in the build container
and records their outputs, losses, gradients, mutated buffers and K-iteration training losses on deterministic
weights/inputs; tests/test_oracle.py checks port.py and train_ref.py against those fixtures on CPU, and np_ops.py
against the torch operators port.py calls.
"""
