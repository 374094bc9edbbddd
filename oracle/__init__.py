"""CPU oracle for the DynamicScaler panorama tiling path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``dynamicscaler_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker /
the timed CPU baseline.

Parity status: PINNED.  The reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the reference itself, imported from
``/root/reference`` in the build container by the committed generators
``tests/golden/make_golden*.py`` and stored as ``tests/golden/*.npz``: every §8a function,
the five BASELINE configs at full size (sha256), ``ddim_step``, ``resize_video_latent``,
``tiled_vae_encode_tensor_simple`` and runs of the REAL denoising loops -- the sphere loops
``basic_sample_shift_shpere_panorama`` of the T2V (``loop.npz``) and I2V (``loop_i2v.npz``) pipelines and the planar loops
``basic_sample_shift_multi_windows`` of the three pipelines that have one (``ring_loop.npz``:
t2v_normal, ``ring_loop_t2v.npz``: t2v_sphere, ``ring_loop_i2v.npz``: i2v_sphere).  ``tests/test_oracle_golden.py`` checks ``panorama_oracle`` and ``torch_port`` against all of them on the CPU.
"""
