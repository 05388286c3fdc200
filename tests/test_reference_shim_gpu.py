"""The reference's own render tests, unchanged, on the B200: tests/ModalRenderTest.cpp + tests/ModalBench.h of khiner/MeshEditor
compiled where they lie and linked against integration/modal_audio_dropin.cpp (the drop-in replacement of src/audio/
ModalAudio.cpp: same entry points, same signatures) and libme_modal.so. oracle/Makefile builds the binary
in the build container (`make -C oracle shim_tests`, part of `ref`); it travels to the GPU box inside the git-ignored
oracle/_ref/. Its three cases (ModalRenderTest.cpp:21-68): excitations superpose linearly (1e-5 of peak), a strike does not
depend on how many threads share it (1e-5 of peak), the click does not depend on the output sample rate (2 %)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

BINARY = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "shim_modal_render_test")


@pytest.mark.skipif(not os.path.exists(BINARY), reason="oracle/_ref/shim_modal_render_test not built (needs /root/reference: make -C oracle shim_tests)")
def test_reference_modal_render_test_passes_through_the_drop_in():
    run = subprocess.run([BINARY], capture_output=True, text=True, timeout=300)
    out = run.stdout + run.stderr
    assert run.returncode == 0, out[-2000:]
    assert "all tests passed" in out and "3 tests" in out, out[-2000:]
