"""CPU logic check of the CUDA sources (no GPU needed): maxwellfdm.jl_b200/csrc is compiled for the host against the
shim of tests/emu (cooperative fibers for CUDA threads, stand-ins for mbarrier / cp.async.bulk, synchronous runtime)
and the resulting TEST-ONLY library is driven through the same C ABI against the oracle.  This is not a product path
and not a fallback (the product binding loads libfdfd_b200.so only; the emulation build says "EMULATED" in
fdfd_version() and nothing times it) - it exists so that indexing, mbarrier-parity, buffer-reuse and missing-wait
mistakes in the kernels are found before a GPU run.  Each group runs in its own process because the ctypes binding
loads one library per process.  See tests/emu/README.md."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))


@pytest.fixture(scope="module")
def emu_lib():
    import build_emu
    return build_emu.build()


def _run(lib, groups, async_mode, shuffle, **extra_env):
    env = dict(os.environ, FDFD_B200_LIB=lib, FDFD_EMU_ASYNC=async_mode, FDFD_EMU_SHUFFLE=str(shuffle), **extra_env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "run_emu_cases.py"), *groups], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, f"{groups} [{async_mode}, shuffle={shuffle}]\n{r.stdout[-2000:]}\n{r.stderr[-4000:]}"
    return r.stdout


# eager: bulk copies land at issue (finds ring stages refilled while still being read);
# lazy: they land at the latest legal moment (finds reads without a barrier wait, staging buffers overwritten before
# the copy engine has read them); shuffle: warps run ahead of / behind each other
MODES = [("eager", 0), ("lazy", 5)]


@pytest.mark.parametrize("mode", MODES, ids=lambda m: f"{m[0]}-shuffle{m[1]}")
def test_emulated_apply_kernels_match_oracle(emu_lib, mode):
    out = _run(emu_lib, ["apply", "boundft", "layouts", "aux"], *mode)
    assert out.count("checks ok") == 4, out


def test_emulated_real_mass_rows(emu_lib):
    """row-pair kernel with real diagonal mass entries streamed as doubles (MDR), a persistent grid of two CTAs so that
    every CTA walks several work items"""
    out = _run(emu_lib, ["realmass"], "eager", 2, FDFD_RP_GRID="2")
    assert "checks ok" in out, out


def test_emulated_fused_full_tensor_shape(emu_lib):
    """fused full-tensor shape of the row-pair kernel (real symmetric off-diagonal rows through the TMA ring): every boundary
    combination, partly empty blocks, three arrangements, transposed apply, fused dots; eager copies + shuffled warps is the
    mode that found the stale per-stage flag after an early release"""
    out = _run(emu_lib, ["fused"], "eager", 7)
    assert "checks ok" in out, out


def test_emulated_single_call_handle(emu_lib):
    """fdfd_multi_* with one slab: the host-side plumbing of csrc/multi.cpp (worker thread, full-grid arrays, layouts)"""
    out = _run(emu_lib, ["multi"], "eager", 0)
    assert "checks ok" in out, out


def test_emulated_row_pair_kernel_without_tensor_maps(emu_lib):
    """the 1-D bulk-copy path of the row-pair kernel (FDFD_RP_TMAP=0; also what the component-major layout takes)"""
    out = _run(emu_lib, ["apply", "deep"], "lazy", 4, FDFD_RP_TMAP="0", FDFD_RP_GRID="3")
    assert out.count("checks ok") == 2, out


def test_emulated_first_generation_kernel(emu_lib):
    """FDFD_K1_GEN=1 keeps the first-generation tiled kernel reachable (A/B timing, on-device cross-check)"""
    out = _run(emu_lib, ["apply", "boundft"], "lazy", 6, FDFD_K1_GEN="1")
    assert out.count("checks ok") == 2, out


def test_emulated_material_pipeline_matches_oracle(emu_lib):
    out = _run(emu_lib, ["matparams"], "eager", 7)
    assert "checks ok" in out, out


def test_emulated_reduced_models_match_k_dimensional_oracle(emu_lib):
    """ModelTE / ModelTM / ModelTEM through the reference call sequence (operator, transpose, right-hand side,
    post-processing, solve) against oracle/reduced.py"""
    out = _run(emu_lib, ["reduced"], "lazy", 2)
    assert "checks ok" in out, out


def test_emulated_multi_chunk_grids_and_offdiag_paths(emu_lib):
    out = _run(emu_lib, ["deep"], "lazy", 3)
    assert "checks ok" in out, out


def test_emulated_correction_pass_skipping_exact_zeros(emu_lib):
    """FDFD_CORR_SKIP_ZERO (opt-in until timed on hardware): threads without off-diagonal material load no field
    values, output cells with zero corner terms are not read-modified-written - same results"""
    out = _run(emu_lib, ["deep"], "eager", 4, FDFD_CORR_SKIP_ZERO="1")
    assert "checks ok" in out, out


def test_emulated_krylov_solvers(emu_lib):
    out = _run(emu_lib, ["solve"], "eager", 0)
    assert "checks ok" in out, out


def _run_dist(world, groups, **modes):
    env = dict(os.environ, **{k: str(v) for k, v in modes.items()})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "run_emu_dist.py"), str(world), *groups],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.count("checks ok") == world, f"{world} {groups} {modes}\n{r.stdout[-2000:]}\n{r.stderr[-4000:]}"


@pytest.mark.parametrize("world,groups,modes", [(3, ["apply"], {}), (2, ["apply"], {"FDFD_PEER_HALO": 0}),
                                                (2, ["apply"], {"FDFD_PEER_HALO": 0, "FDFD_HALO_OVERLAP": 1}),
                                                (3, ["krylov"], {"FDFD_PEER_DIRECT": 1}),
                                                (2, ["apply"], {"FDFD_SPLIT_OVERLAP": 1}),
                                                (2, ["apply"], {"FDFD_PEER_HALO": 0, "FDFD_INKERNEL_HALO_WAIT": 1, "FDFD_K1_GEN": 1})],
                         ids=["default-peer-overlap-3", "nccl-2", "nccl-overlap-2", "peer-direct-3", "split-overlap-2",
                              "gen1-inkernel-wait-2"])
def test_emulated_z_slab_ranks(emu_lib, world, groups, modes):
    """one process per rank as on the GPU box; NCCL / driver entry points replaced by tests/emu/fakelibs, device memory
    in named shared memory so that the CUDA-IPC peer-halo path (the default data plane) maps between the processes; NCCL
    send/recv forced with FDFD_PEER_HALO=0, the in-kernel halo wait on either plane: halo exchange on every
    arrangement and layout, Bloch wrap between the first and last rank, back-to-back applies (protocol epochs), host
    and device paths, transposed apply, BiCGSTAB / QMR with allreduced dots"""
    _run_dist(world, groups, **modes)


def test_product_binding_never_points_at_the_emulation_build():
    """the default library path of the product binding is the CUDA build; the emulation library is only reachable
    through the explicit FDFD_B200_LIB override used above"""
    import maxwellfdm_jl_b200 as fb
    if "FDFD_B200_LIB" not in os.environ:
        assert fb._lib.LIB_PATH.endswith(os.path.join("maxwellfdm.jl_b200", "libfdfd_b200.so"))
        assert b"sm_100a" in fb._lib.lib().fdfd_version()
