"""CPU oracle for the UAPS unlabeled-batch hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``uaps_b200/`` imports this package; it
is imported by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs as the *checker* and the timed
CPU baseline, never as the product path.

Pinning.  The reference (djene-mengistu/UAPS) ships no tests, no golden
vectors and no weights, so there is nothing of the
reference's own to pin against.  Instead every function here is checked, in
the build container, against the *imported reference modules themselves*
(``/root/reference/utilities/{UAPS_unet,pytorch_losses,ramps}.py``) by
``oracle/make_golden.py``, which also writes the fixtures under
``tests/golden/``.  The inline loss section of ``UAPS_train.py`` cannot be
imported (it is straight-line code in a loop body behind missing
dependencies); it is restated expression by expression in
``oracle/uaps_loss_ref.py`` with the reference line cited beside each one.
Parity for that section is therefore "pinned to the reference's imported
building blocks (dice_loss, torch ops), restated for the inline glue".

All arithmetic is PyTorch (torch 2.11.0+cu128 in this image) -- the reference
states only "PyTorch >= 1.1.0" (README.md:44-46) and pins nothing.
"""
