"""CPU oracle for the chunk-encoder + feature-alignment hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``large-scale-vit-slam_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may.  It is the checker, never the product.

What it is
----------
* ``oracle.functional``  – fp32 PyTorch restatement of the UPSTREAM ``facebookresearch/vggt``
  symbols the reference imports but does not vendor (Aggregator, DINOv2 ViT-L/14-reg, Block,
  Attention, Mlp, LayerScale, 2-D RoPE, CameraHead, pose_enc, rotation, closed_form_inverse_se3).
  Upstream is pinned nowhere by the reference (README.md:45-46 clones HEAD) and is absent from
  this container, so this part is restated from the published algorithm:
  **parity for rows a1–a5 is "vs. our restatement", i.e. unpinned by upstream itself.**
* ``oracle.aligned``     – fp32 restatement of the reference-owned code on the path
  (alignment head, cross attention, 1-D RoPE, gated update, pose chain, Sim(3) apply, IRLS
  Umeyama, chunk generation).  This part IS pinned: ``oracle/make_golden.py`` imports the real
  files from ``/root/reference`` (on top of ``oracle/vggt_shim``) in the build container,
  checks the restatement against them and writes ``tests/golden/*.npz``.
* ``oracle/vggt_shim``   – importable ``vggt.vggt.*`` namespace (thin nn.Module containers that
  delegate to ``oracle.functional``) so the reference's own files import unmodified.

The reference has no tests / golden vectors of its own, hence the fixtures are
generated from the reference's code run here, with the generating script committed.
"""
