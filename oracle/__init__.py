"""CPU oracles for the tiled full-DEM super-resolution path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it,
and there only as the checker / the CPU arm, never as the thing measured or shipped.

* ``oracle.tiling``     -- numpy restatement of the reference's pad / tile / patch / blend / assemble code
                           (``process_full_tiles.py:246-587``).  PINNED: checked against outputs of the
                           reference's own ``DEMSuperResolution`` class run in the build container
                           (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).
* ``oracle.generator``  -- torch-CPU restatement of the SPADE / CNN-SPADE / pix2pix generators
                           (``spade/models/*.py``, ``pix2pix.py:64-108``).  PARITY UNPINNED: the reference's
                           arithmetic lives in TensorFlow 2.5 / tensorflow-addons 0.16.1, neither is installed and
                           the reference ships no golden vectors.
"""
