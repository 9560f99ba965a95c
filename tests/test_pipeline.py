"""CPU: the generation / post-processing overlap (pipeline.GenerationPipeline) with stand-ins for the GPU generator."""
import threading
import time

import numpy as np
import pytest

from ml_conformer_generator_b200.pipeline import GenerationPipeline, sdf_postprocess


def _fake_generator(B, N, gen_s, log):
    def generate(k, out):
        if out is None:
            out = (np.zeros((B, N, 3), np.float32), np.zeros((B, N), np.int32), np.zeros((B, 42, 42), np.int8),
                   np.zeros(B, np.int32))
        x, cls, bonds, nn = out
        log.append(("gen_start", k, time.perf_counter()))
        time.sleep(gen_s)                      # the GPU call blocks with the GIL released
        x[:] = k                               # overwrites the ring slot in place, like pinned output buffers
        cls[:] = k % 7
        bonds[:] = 0
        nn[:] = 5 + (k % 3)
        log.append(("gen_end", k, time.perf_counter()))
        return x, cls, bonds, nn
    return generate


def test_overlap_order_and_buffer_reuse():
    B, N, n_batches, gen_s, post_s = 32, 8, 6, 0.08, 0.004
    log = []

    def post(x, cls, bonds, n, index):
        k = index // B
        assert n == 5 + (k % 3) and x.shape == (n, 3)
        assert float(x[0, 0]) == float(k) and int(cls[0]) == k % 7   # the slot still holds batch k, not k + depth
        time.sleep(post_s)
        return None if index % 4 == 3 else index                       # every 4th molecule "fails sanitisation"

    pipe = GenerationPipeline(_fake_generator(B, N, gen_s, log), post, n_workers=4, depth=2, chunk=8)
    t0 = time.perf_counter()
    out = list(pipe.run(n_batches))
    wall = time.perf_counter() - t0
    assert len(out) == n_batches
    for k, batch in enumerate(out):
        assert batch == [i for i in range(k * B, (k + 1) * B) if i % 4 != 3]    # in order, failures dropped
    assert pipe.stats == {"batches": n_batches, "molecules": n_batches * B, "kept": n_batches * B * 3 // 4}
    serial = n_batches * (gen_s + B * post_s / 4)
    assert wall < 0.85 * serial, (wall, serial)       # post-processing of batch k hid behind generation of batch k + 1
    # the generator never ran more than `depth` batches ahead of the consumer
    starts = {k: t for ev, k, t in log if ev == "gen_start"}
    assert all(starts[k + 1] >= starts[k] for k in range(n_batches - 1))


def test_generator_error_is_raised_in_the_consumer():
    def generate(k, out):
        if k == 1:
            raise ValueError("boom")
        return (np.zeros((2, 4, 3), np.float32), np.zeros((2, 4), np.int32), np.zeros((2, 42, 42), np.int8), np.full(2, 4, np.int32))

    pipe = GenerationPipeline(generate, lambda *a: 1, n_workers=2)
    it = pipe.run(3)
    assert next(it) == [1, 1]
    with pytest.raises(ValueError):
        next(it)
    with pytest.raises(ValueError):
        GenerationPipeline(generate, lambda *a: 1, depth=1)


def test_default_postprocess_writes_a_mol_block():
    x = np.array([[0.0, 0.0, 0.0], [1.5, 0.0, 0.0], [2.2, 1.2, 0.0]], np.float32)
    cls = np.array([0, 0, 2], np.int32)
    bonds = np.zeros((42, 42), np.int8)
    bonds[1, 0] = 1
    bonds[2, 1] = 2
    block = sdf_postprocess(x, cls, bonds, 3, 17)
    lines = block.splitlines()
    assert lines[0] == "mlcg_17" and "PROVISIONAL" in lines[2]
    assert lines[3].startswith("  3  2") and lines[-2] == "M  END" and lines[-1] == "$$$$"
