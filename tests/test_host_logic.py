"""CPU tests of the host side: C-ABI surface (library loads, every symbol of include/otb.h is exported, struct
sizes agree with the compiler), scene flattening, ray sharding and the gloo world-size-2 collectives, and the
Python -> device-function translator (checked by compiling the generated code for the host with g++)."""
import ctypes as C
import os
import pathlib
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import optrace_b200 as ot
from optrace_b200 import _cabi, dist, userfunc
from optrace_b200.scene import flatten_raytracer, detector_record
import scenes

ROOT = pathlib.Path(__file__).resolve().parent.parent


# ---- C ABI -----------------------------------------------------------------------------------------------
def _header_symbols():
    txt = (ROOT / "include" / "otb.h").read_text()
    return sorted(set(re.findall(r"^(?:int|const char\*)\s+(otb_\w+)\s*\(", txt, flags=re.M)))


def test_library_exports_every_header_symbol():
    lib = _cabi.lib()            # raises if libotb.so is missing: there is no CPU fallback to hide behind
    names = _header_symbols()
    assert len(names) >= 20
    assert sorted(_cabi.SYMBOLS) == names
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.otb_abi_version() == 2


def test_struct_layouts_match_the_compiler():
    """sizeof of every ABI struct as seen by gcc equals the ctypes mirror"""
    structs = ["OtbSurface", "OtbMedium", "OtbFilter", "OtbStep", "OtbSceneDesc", "OtbRays", "OtbRayStore",
               "OtbDetector", "OtbSource", "OtbDeviceInfo"]
    prog = '#include <stdio.h>\n#include "otb.h"\nint main(){' + "".join(
        f'printf("%zu\\n", sizeof({s}));' for s in structs) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        src = pathlib.Path(d) / "s.c"
        src.write_text(prog)
        subprocess.run(["gcc", f"-I{ROOT / 'include'}", str(src), "-o", f"{d}/s"], check=True)
        sizes = [int(x) for x in subprocess.run([f"{d}/s"], capture_output=True, text=True, check=True).stdout.split()]
    for s, n in zip(structs, sizes):
        assert C.sizeof(getattr(_cabi, s)) == n, s


def test_engine_refuses_to_run_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    RT = scenes.spherical_aberration(ot)
    with pytest.raises(_cabi.EngineError):
        RT.trace(1000)
    assert _cabi.lib().otb_init(0) != 0      # the C entry point fails loudly as well


# ---- flattening ----------------------------------------------------------------------------------------------
def test_flatten_double_gauss():
    RT = scenes.double_gauss(ot)
    fs = flatten_raytracer(RT)
    assert fs.nt == 17 and len(fs.steps) == 16
    roles = [s["role"] for s in fs.steps]
    assert roles == [0, 1, 0, 1, 0, 1, 4, 0, 1, 0, 1, 0, 1, 0, 1, 4]
    assert fs.steps[-1]["hurb"] == 0 and fs.surfaces[fs.steps[-1]["surface"]]["kind"] == 1
    # media are de-duplicated by value: ambient + 5 distinct Abbe glasses (n=1.773 and n=1.788 are each used twice)
    assert len(fs.media) == 6
    d = fs.to_ctypes()
    assert d.n_steps == 16 and d.n_surfaces == 16
    assert fs.fingerprint() == flatten_raytracer(scenes.double_gauss(ot)).fingerprint()


def test_flatten_hurb_and_filters():
    fs = flatten_raytracer(scenes.hurb_aperture(ot, "Pinhole"))
    assert fs.n_hurb == 1 and [s["hurb"] for s in fs.steps] == [1, 0]
    fs = flatten_raytracer(scenes.zoo_analytic(ot))
    assert len(fs.filters) == 4 and {f["type"] for f in fs.filters} == {0, 1, 2, 3}
    assert fs.aux.shape[0] > 0


def test_snapshot_detects_changes():
    RT = scenes.spherical_aberration(ot)
    a = RT.tracing_snapshot()
    RT.detectors[0].move_to([0, 0, 30])            # detectors do not influence the trace
    assert RT.tracing_snapshot() == a
    RT.lenses[0].move_to([0, 0, 1])
    assert RT.tracing_snapshot() != a


def test_api_errors_like_the_reference():
    RT = scenes.spherical_aberration(ot)
    with pytest.raises(TypeError):
        RT.trace(1000.0)
    with pytest.raises(ValueError):
        RT.trace(0)
    with pytest.raises(RuntimeError):
        RT.detector_image()                          # no rays traced
    with pytest.raises(ValueError):
        ot.Raytracer(outline=[0, 1, 0, 1, 1, 0])
    with pytest.raises(ValueError):
        ot.ConicSurface(r=5, R=3, k=1)               # r beyond the conic section
    with pytest.raises(ValueError):
        ot.RingSurface(r=1, ri=2)
    with pytest.raises(RuntimeError):
        ot.Detector(ot.FunctionSurface2D(r=1, func=lambda x, y: 0*x), pos=[0, 0, 0])


def test_ray_split_and_storage_size():
    from optrace_b200.ray_storage import RayStorage, split_rays
    np.random.seed(0)
    n = split_rays(1001, [1.0, 2.0, 1.0])
    assert n.sum() == 1001 and abs(n[1] - 500) <= 2
    assert RayStorage.storage_size(1000, 17, False) == 1000*(17*48 + 28)
    assert RayStorage.storage_size(1000, 17, True) == 1000*(17*36 + 28) + 8
    assert RayStorage.max_rays_for_size(RayStorage.storage_size(1000, 17, False), 17, False) == 1000


# ---- sharding + collectives (gloo, world size 2) -----------------------------------------------------------------
def test_shard_ranges_cover_all_rays():
    for N in (10, 1001, 10_000_000):
        for G in (1, 2, 3, 8):
            r = [dist.shard_range(N, k, G) for k in range(G)]
            assert r[0][0] == 0 and r[-1][1] == N
            assert all(r[k][1] == r[k + 1][0] for k in range(G - 1))
    sl = dist.source_slices([0, 40, 100], 30, 70)
    assert sl == [(0, 0, 10), (1, 10, 30)]


def _gloo_worker(rank, world, port, out):
    import torch
    import torch.distributed as td
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    from optrace_b200 import dist as d
    assert d.is_dist() and d.world() == world and d.rank() == rank
    b, e = d.shard_range(1001)
    img = torch.full((4, 3, 4), float(rank + 1), dtype=torch.float64)
    d.allreduce_sum_(img)
    rng = torch.tensor([float(rank), float(rank) + 1, -float(rank), 5.0 - rank], dtype=torch.float64)
    d.allreduce_range_(rng)
    msgs = torch.tensor([b, e], dtype=torch.int64)
    d.allreduce_sum_(msgs)
    nl = d.broadcast_ints(np.array([7 + rank, 3]), torch.device("cpu"))
    out.put((rank, (b, e), float(img[0, 0, 0]), rng.tolist(), msgs.tolist(), nl.tolist()))
    td.destroy_process_group()


def test_collectives_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in ps]
    assert [r[1] for r in res] == [(0, 500), (500, 1001)]
    for r in res:
        assert r[2] == 3.0                              # 1 + 2
        assert r[3] == [0.0, 2.0, -1.0, 5.0]            # min x, max x, min y, max y over ranks
        assert r[4] == [500, 1501]
        assert r[5] == [7, 3]                           # rank 0's counts everywhere


# ---- user-callable translator ----------------------------------------------------------------------------------
def _compile_host(header: str):
    d = tempfile.mkdtemp()
    src = pathlib.Path(d) / "u.cpp"
    src.write_text('#include <math.h>\n#define __device__\n#define __forceinline__ inline\n' + header +
                   '\nextern "C" double f1(int id, double a){return otb_user_f1(id,a);}\n'
                   'extern "C" double f2(int id, double a, double b){return otb_user_f2(id,a,b);}\n'
                   'extern "C" void d2(int id, double a, double b, double* x, double* y){otb_user_d2(id,a,b,x,y);}\n')
    so = pathlib.Path(d) / "u.so"
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", str(src), "-o", str(so)], check=True)
    l = C.CDLL(str(so))
    l.f1.restype = l.f2.restype = C.c_double
    l.f1.argtypes = [C.c_int, C.c_double]
    l.f2.argtypes = [C.c_int, C.c_double, C.c_double]
    l.d2.argtypes = [C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    return l


K = 0.37


def _g2(x, y, a=2.0):
    t = np.sqrt(x**2 + y**2)
    u = np.where(t < 1.5, np.cos(a*t)*K, np.exp(-t))
    return u + np.abs(x)*np.arctan2(y, x + 5) - y**3/7


def test_translator_matches_numpy():
    fs = flatten_raytracer(scenes.zoo_numeric(ot))
    funcs = list(fs.user_funcs) + [("surf2d", _g2, dict(a=1.3)), ("surf2d", lambda x, y: 0.1*np.cos(2*np.pi*x/2), {}),
                                   ("mask2d", lambda x, y: (x > -1) & (np.hypot(x, y) <= 2.5), {})]
    lib = _compile_host(userfunc.generate_header(funcs))
    rng = np.random.default_rng(0)
    x, y = rng.uniform(-3, 3, 200), rng.uniform(-3, 3, 200)
    for i, (kind, fn, kw) in enumerate(funcs):
        if kind == "deriv2d":
            ex, ey = fn(x, y, **kw)
            a, b = C.c_double(), C.c_double()
            for j in range(x.size):
                lib.d2(i, x[j], y[j], C.byref(a), C.byref(b))
                assert abs(a.value - ex[j]) <= 1e-15*max(1, abs(ex[j])) and abs(b.value - ey[j]) <= 1e-15*max(1, abs(ey[j]))
        elif kind.endswith("1d"):
            ref = np.asarray(fn(np.abs(x), **kw), dtype=np.float64)
            got = np.array([lib.f1(i, v) for v in np.abs(x)])
            assert np.allclose(got, ref, rtol=1e-15, atol=0), kind
        else:
            ref = np.asarray(fn(x, y, **kw), dtype=np.float64)
            got = np.array([lib.f2(i, a, b) for a, b in zip(x, y)])
            assert np.allclose(got, ref, rtol=2e-15, atol=1e-16), kind


def test_translator_rejects_unsupported_code():
    def bad(x, y):
        return np.fft.fft(x)
    with pytest.raises(NotImplementedError):
        userfunc.translate("surf2d", bad, {}, "f")


def test_stratum_permutation_is_a_bijection(tmp_path):
    """otb_rng.cuh feistel_perm (the stand-in for the reference's shuffle of stratified samples, random.py:41-45)
    is host-callable: compiled with nvcc as host code it must map [0, n) onto itself for awkward n as well"""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not pathlib.Path(nvcc).exists():
        pytest.skip("nvcc not available")
    src = tmp_path / "perm.cu"
    src.write_text(r'''
#include <cstdio>
#include <vector>
#include <cmath>
#include "otb_rng.cuh"
int main() {
    const unsigned long long ns[] = {2, 3, 5, 16, 17, 100, 1000, 4097, 65536, 65537, 1000003};
    for (unsigned long long n : ns) {
        std::vector<char> seen(n, 0);
        double corr = 0;
        for (unsigned long long i = 0; i < n; ++i) {
            unsigned long long p = feistel_perm(i, n, 0x1234567890abcdefull);
            if (p >= n || seen[p]) { printf("FAIL %llu\n", n); return 1; }
            seen[p] = 1;
            corr += (double)i*(double)p;
        }
        double mean = (n - 1)/2.0, var = ((double)n*n - 1)/12.0;
        if (n > 1000 && std::fabs((corr/n - mean*mean)/var) > 0.05) { printf("CORR %llu\n", n); return 2; }
    }
    printf("OK\n");
    return 0;
}
''')
    exe = tmp_path / "perm"
    r = subprocess.run([nvcc, "-x", "cu", "-I", str(ROOT / "optrace_b200" / "csrc"), "-o", str(exe), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "OK", out.stdout


def test_balanced_source_shards_tile_every_source():
    """dist.shard_sources: every rank holds a slice of every source, the slices of all ranks tile the source's block of
    global ray ids, local storage rows map back to the right source (RayStorage._local_range)"""
    from optrace_b200 import dist
    from optrace_b200.ray_storage import RayStorage
    N_list = [1000, 1, 0, 777, 12]
    B = np.concatenate(([0], np.cumsum(N_list)))
    for G in (1, 2, 3, 8):
        seen = np.zeros(int(B[-1]), dtype=int)
        for r in range(G):
            blocks = dist.shard_sources(N_list, r, G)
            assert [i for i, _, _ in blocks] == sorted(i for i, _, _ in blocks)
            for i, g0, c in blocks:
                assert B[i] <= g0 and g0 + c <= B[i + 1] and c > 0
                seen[g0:g0 + c] += 1

            class _Store:
                N, nt = sum(c for _, _, c in blocks), 3
            rs = RayStorage()
            rs._attach(_Store(), [None]*len(N_list), N_list, False, int(B[-1]), blocks)
            off = 0
            for i, g0, c in blocks:
                assert rs._local_range(i) == (off, off + c)
                off += c
            for i in set(range(len(N_list))) - {i for i, _, _ in blocks}:
                b, e = rs._local_range(i)
                assert b == e
        assert np.all(seen == 1)
    # a contiguous range given as its first global id gives the same blocks as before
    rs = RayStorage()

    class _S2:
        N, nt = 40, 2
    rs._attach(_S2(), [None]*3, [40, 60, 10], False, 110, 30)
    assert rs._blocks == [(0, 30, 10), (1, 40, 30)] and rs.ray_begin == 30 and rs._local_range(1) == (10, 40)


def test_reference_gaussian_filter_weights_depend_on_the_hosts_simd_level():
    """Justifies golden_util.W_RTOL (deviation 1).  The reference evaluates a Gaussian
    TransmissionSpectrum with numpy's FLOAT32 exp (spectrum.py:113 + NEP 50).  That loop is dispatched by CPU feature:
    the AVX2/AVX512 kernels are accurate to ~2.5 ulp, the baseline loop calls libm's (practically correctly rounded)
    expf.  The same reference therefore produces weights that differ by up to 2 float32 ulp (1.9e-7) between hosts,
    so no engine can match it to 1e-9 on every host; the engine rounds a float64 exp once, which is what the
    baseline loop gives in > 99.9 % of the arguments."""
    import os
    import subprocess
    import sys
    import numpy as np
    try:
        from numpy._core._multiarray_umath import __cpu_features__ as feats
    except Exception:
        pytest.skip("numpy CPU feature table not available")
    simd = [k for k, v in feats.items() if v and (k.startswith("AVX") or k == "FMA3")]
    if "AVX2" not in simd:
        pytest.skip("this host already runs numpy's baseline float32 exp")
    code = ("import numpy as np, sys; x = -(np.linspace(0, 30, 200001).astype(np.float32)); "
            "sys.stdout.buffer.write(np.exp(x).tobytes())")
    env = dict(os.environ, NPY_DISABLE_CPU_FEATURES=" ".join(simd))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, check=True).stdout
    base = np.frombuffer(out, dtype=np.float32)                       # numpy on a host without AVX
    x = -(np.linspace(0, 30, 200001).astype(np.float32))
    here = np.exp(x)                                                  # numpy on this host (SIMD loop)
    engine_rule = np.exp(x.astype(np.float64)).astype(np.float32)    # float64 exp rounded once (otb_media.cuh)
    rel = np.abs(here.astype(np.float64) - base)/base
    assert (here != base).mean() > 0.05            # the reference's own numbers differ between the two hosts ...
    assert 5e-8 < rel.max() < 3e-7                 # ... by up to ~2 float32 ulp: W_RTOL = 3e-7 covers exactly that
    assert (engine_rule != base).mean() < 1e-2     # the engine's rule IS the baseline loop up to rare last-bit ties


def test_shared_split_cache_hands_out_copies():
    """dist.shared_split caches the deterministic split of (N, powers); callers may modify what they get"""
    import numpy as np
    from optrace_b200 import dist
    from optrace_b200.ray_storage import split_rays
    a = dist.shared_split(1000, [1.0, 3.0], None)
    assert a.tolist() == split_rays(1000, [1.0, 3.0]).tolist() == [250, 750]
    a[0] = -1
    b = dist.shared_split(1000, [1.0, 3.0], None)
    assert b.tolist() == [250, 750]
    assert dist.shared_split(1000, [3.0, 1.0], None).tolist() == [750, 250]      # another key, not the cached one
