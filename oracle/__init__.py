"""CPU oracle for the DispNet hot path -- TEST INFRASTRUCTURE ONLY.

This package is a torch-fp32 / numpy restatement of the arithmetic the reference
(zenithfang/supervised_dispnet @ f81dfcc) performs on its training hot path.  Every function cites the
reference file:line it follows.  It exists to *check* the CUDA product path, never to *be* it:

  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
    legs may import anything from here;
  * nothing under ``supervised_dispnet_b200/`` imports it (tests/test_boundary.py greps for that).

Parity pin: the reference ships no golden vectors, KATs or fixtures, so the oracle is
pinned against outputs of the reference *itself*, generated in the authoring container by
``oracle/make_golden.py`` (which imports ``/root/reference``) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` replays them on CPU.

The arithmetic of the path lives in a third-party dependency that is absent from ``/root/reference``:
PyTorch (pinned ``torch==1.0.1`` in requirements.txt:1; executed here with torch 2.11.0) and torchvision's
``vgg16_bn`` topology.  The oracle therefore restates the *published* semantics of those ops (conv as
cross-correlation, BatchNorm2d training statistics, bilinear grid_sample with ``align_corners=False`` as
executed by the container's torch -- SURVEY.md hard part 6) and anchors on the reference's call sites.
"""
