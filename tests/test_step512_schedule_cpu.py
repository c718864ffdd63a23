"""The 512x512 training step (BASELINE configs[4] shapes) on the CPU over the kernel emulations
(tests/kernel_emulators.py): the host-side schedule of tests/test_step512_gpu.py — 19 AdaIN sites, 8 up-blocks,
7 discriminator blocks on a 512x512 input, all six criteria, both backward passes into the gradient sinks — against
the same golden of the unmodified reference (tests/golden/step512.pt).  Not kernel evidence: the kernels are emulated
with plain torch ops here; the GPU test runs the identical body with the real ones."""
import kernel_emulators as E
import test_step512_gpu as T


def test_512_step_schedule_on_emulated_kernels(monkeypatch):
    E.install(monkeypatch)
    monkeypatch.setattr(T, "DEV", "cpu")
    monkeypatch.setattr(T, "REPORT_NAME", "step512_gradient_errors_cpu_emulation.json")
    T.test_512_training_step()
