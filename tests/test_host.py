"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the FFT cores are right,
plan geometry errors are reported, and the Protocol surface exists.  No compute calls (there is no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from modular_audio_pipeline_b200 import _capi
    lib = _capi.lib()
    header = open(os.path.join(ROOT, "include", "maf.h")).read()
    declared = set(re.findall(r"\b(maf_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/maf.h but not exported"
    assert declared == set(_capi.SYMBOLS), "ctypes table and header disagree"
    assert lib.maf_version() == 1


def test_calls_fail_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from modular_audio_pipeline_b200 import _capi, FrontEnd
    h = C.c_void_p()
    assert _capi.lib().maf_create(0, C.byref(h)) == _capi.MAF_ERR_CUDA
    assert b"cuda" in _capi.lib().maf_last_error().lower()
    with pytest.raises(RuntimeError):
        FrontEnd()


def test_fft_cores_on_cpu(tmp_path):
    exe = tmp_path / "host_fft_check"
    subprocess.check_call(["g++", "-O2", "-o", str(exe), os.path.join(ROOT, "tests", "host", "host_fft_check.cpp")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "OK" in out.stdout


# ---- WAV container I/O (wavio.py; reference read_wave / write_wave, preprocessor.py:71-90) ---------------------------
def _wave_write(path, raw, channels, width, rate):
    import wave
    with wave.open(str(path), "wb") as wf:
        wf.setnchannels(channels); wf.setsampwidth(width); wf.setframerate(rate); wf.writeframes(raw)


def test_wavio_reader_matches_wave_module(tmp_path):
    import struct
    import wave
    import numpy as np
    import torch
    from modular_audio_pipeline_b200 import wavio
    rng = np.random.default_rng(0)
    pool = wavio.PinnedPool(pinned=False)
    for i, (ch, width, rate, nfr) in enumerate([(1, 2, 16000, 12345), (2, 2, 48000, 4001), (1, 2, 8000, 0), (1, 1, 22050, 778),
                                                (2, 1, 44100, 100), (1, 2, 16000, 1)]):
        raw = rng.integers(0, 256, size=nfr * ch * width, dtype=np.uint8).tobytes()
        p = tmp_path / f"w{i}.wav"
        _wave_write(p, raw, ch, width, rate)
        with wave.open(str(p), "rb") as wf:
            want = wf.readframes(wf.getnframes())
        x, info = wavio.read_pcm16(str(p), pool)
        assert (info.channels, info.sampwidth, info.rate, info.nframes) == (ch, width, rate, nfr)
        assert x.numpy().tobytes() == want == raw
    # chunks before / after the data chunk, odd-sized chunk padding, WAVE_FORMAT_EXTENSIBLE
    pcm = rng.integers(-32768, 32767, size=999, dtype=np.int16).tobytes()
    fmt = struct.pack("<HHIIHH", 1, 1, 16000, 32000, 2, 16)
    ext = struct.pack("<HHIIHHHHI", 0xFFFE, 1, 16000, 32000, 2, 16, 22, 16, 4) + b"\x01\x00\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"
    for j, f in enumerate((fmt, ext)):
        body = (b"WAVE" + b"LIST" + struct.pack("<I", 5) + b"abcde\x00" + b"fmt " + struct.pack("<I", len(f)) + f +
                b"data" + struct.pack("<I", len(pcm)) + pcm + b"cue " + struct.pack("<I", 4) + b"\x00" * 4)
        p = tmp_path / f"x{j}.wav"
        p.write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)
        with wave.open(str(p), "rb") as wf:
            want = wf.readframes(wf.getnframes())
        x, info = wavio.read_pcm16(str(p))
        assert x.numpy().tobytes() == want == pcm and info.rate == 16000
    # not PCM: the wave module's own error comes through
    bad = tmp_path / "bad.wav"
    body = b"WAVE" + b"fmt " + struct.pack("<I", 16) + struct.pack("<HHIIHH", 3, 1, 16000, 64000, 4, 32) + b"data" + struct.pack("<I", 8) + b"\x00" * 8
    bad.write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)
    with pytest.raises(wave.Error):
        wavio.read_pcm16(str(bad))
    with pytest.raises(Exception):
        wavio.read_pcm16(str(tmp_path / "missing.wav"))


def test_wavio_writer_is_byte_identical_to_wave_module(tmp_path):
    import numpy as np
    import torch
    from modular_audio_pipeline_b200 import wavio
    rng = np.random.default_rng(1)
    for n, rate in ((0, 16000), (1, 8000), (16001, 16000), (48000, 48000)):
        x = torch.from_numpy(rng.integers(-32768, 32767, size=n, dtype=np.int16))
        a, b = tmp_path / "a.wav", tmp_path / "b.wav"
        wavio.write_pcm16(str(a), x, rate)
        _wave_write(b, x.numpy().tobytes(), 1, 2, rate)
        assert a.read_bytes() == b.read_bytes()


def test_wav_prefetcher_order_and_buffer_lifetime(tmp_path):
    import numpy as np
    import torch
    from modular_audio_pipeline_b200 import wavio
    rng = np.random.default_rng(2)
    paths, want = [], []
    for i in range(7):
        x = rng.integers(-32768, 32767, size=1000 + 317 * i, dtype=np.int16)
        p = tmp_path / f"p{i}.wav"
        _wave_write(p, x.tobytes(), 1, 2, 16000)
        paths.append(str(p)); want.append(x)
    for depth in (2, 3):
        held = []
        for i, (p, x, info) in enumerate(wavio.WavPrefetcher(paths, depth=depth, pinned=False)):
            assert p == paths[i] and info.rate == 16000 and np.array_equal(x.numpy(), want[i])
            held.append((i, x))
            held = held[-(depth - 1):]                      # items still guaranteed valid
            for j, xj in held:
                assert np.array_equal(xj.numpy(), want[j])
    with pytest.raises(Exception):
        list(wavio.WavPrefetcher([paths[0], str(tmp_path / "nope.wav")], pinned=False))


def test_run_merge_logic_on_cpu(tmp_path):
    """csrc/run_merge.h (per-thread run summaries + associative merge, as used by select_clip_kernel and the grid-wide
    selection) against the literal loop of the reference's _detect_noise_segments, randomised, incl. empty and all-noise."""
    exe = tmp_path / "host_runs_check"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "host", "host_runs_check.cpp")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "OK" in out.stdout


def test_ctypes_table_matches_header_arity_and_pointer_positions():
    """Every prototype in include/maf.h against the ctypes table: same number of parameters, pointers where the header has
    pointers, 64-bit integers where it has int64_t / size_t / long long, double / float where it says so."""
    from modular_audio_pipeline_b200 import _capi
    header = open(os.path.join(ROOT, "include", "maf.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    protos = re.findall(r"\b(?:const\s+char\s*\*|int64_t|int|void|size_t|long\s+long)\s+(maf_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S)
    assert len(protos) >= 30
    seen = set()
    for name, params in protos:
        seen.add(name)
        params = " ".join(params.split())
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        res, args = _capi.SYMBOLS[name]
        assert len(plist) == len(args), f"{name}: header has {len(plist)} parameters, ctypes table {len(args)}"
        for p, a in zip(plist, args):
            if "*" in p:
                assert a not in (C.c_int, C.c_int32, C.c_int64, C.c_size_t, C.c_float, C.c_double), f"{name}: '{p}' is a pointer, table has {a}"
            elif re.search(r"\b(int64_t|size_t|long long)\b", p):
                assert a in (C.c_int64, C.c_size_t, C.c_longlong), f"{name}: '{p}' vs {a}"
            elif re.search(r"\bdouble\b", p):
                assert a is C.c_double, f"{name}: '{p}' vs {a}"
            elif re.search(r"\bfloat\b", p):
                assert a is C.c_float, f"{name}: '{p}' vs {a}"
            elif re.search(r"\b(int32_t|int)\b", p):
                assert a in (C.c_int32, C.c_int), f"{name}: '{p}' vs {a}"
            else:
                raise AssertionError(f"{name}: unrecognised parameter '{p}'")
    assert seen == set(_capi.SYMBOLS), sorted(set(_capi.SYMBOLS) ^ seen)


def test_ctypes_structs_match_header_layout(tmp_path):
    """sizeof / offsetof of the three structs in include/maf.h as the C compiler lays them out, against the ctypes mirrors."""
    from modular_audio_pipeline_b200 import _capi
    structs = {"maf_item_desc": _capi.ItemDesc, "maf_params": _capi.Params, "maf_layout": _capi.Layout}
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "maf.h"', 'int main(void) {']
    for cname, cls in structs.items():
        src.append(f'printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            src.append(f'printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    src += ['return 0; }']
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(c)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    seen = 0
    for line in out.splitlines():
        cname, field, val = line.split()
        cls = structs[cname]
        if field == "size":
            assert C.sizeof(cls) == int(val), line
        else:
            assert getattr(cls, field).offset == int(val), line
        seen += 1
    assert seen == sum(len(cls._fields_) + 1 for cls in structs.values())


def test_property_wavio_roundtrip_and_wave_module_interop(tmp_path):
    """hypothesis: write_pcm16 -> read_pcm16 is the identity, the wave module reads our files and we read the wave module's,
    for arbitrary lengths (incl. 0 and odd), rates and contents."""
    import wave
    import numpy as np
    import torch
    pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st
    from modular_audio_pipeline_b200 import wavio
    pool = wavio.PinnedPool(pinned=False)

    @settings(max_examples=60, deadline=None)
    @given(n=st.integers(0, 5000), rate=st.sampled_from([8000, 16000, 22050, 44100, 48000]), seed=st.integers(0, 2 ** 31 - 1))
    def check(n, rate, seed):
        x = torch.from_numpy(np.random.default_rng(seed).integers(-32768, 32768, size=n).astype(np.int16))
        p = str(tmp_path / "rt.wav")
        wavio.write_pcm16(p, x, rate)
        y, info = wavio.read_pcm16(p, pool)
        assert torch.equal(x, y) and (info.rate, info.channels, info.sampwidth, info.nframes) == (rate, 1, 2, n)
        with wave.open(p, "rb") as wf:
            assert (wf.getframerate(), wf.getnchannels(), wf.getsampwidth(), wf.getnframes()) == (rate, 1, 2, n)
            assert wf.readframes(n) == x.numpy().tobytes()
        i2 = wavio.read_info(p)
        out = torch.empty(n, dtype=torch.int16)
        wavio.read_data_into(p, i2, out)
        assert torch.equal(out, x)

    check()


# ---- the bench line contract, checked on the committed record of the last full run ---------------------------------------------
def test_committed_bench_line_has_the_contract_keys():
    """`python bench.py` prints ONE JSON line; the driver and the judge read these keys.  The committed record of the round's last
    full run (profiles/r02_final_bench_cfg2_n1.json) must carry every one of them with sane values."""
    import json
    rec = json.load(open(os.path.join(ROOT, "profiles", "r02_final_bench_cfg2_n1.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in rec, k
    assert rec["metric"] == "audio_seconds_per_second" and rec["unit"] == "audio-s/s" and rec["higher_is_better"] is True
    assert rec["scaling"] == "weak" and rec["vs_baseline"] is None and rec["data"] == "synthetic" and "workload" in rec["config"]
    assert abs(rec["value"] - 3600.0 * rec["n_gpus"] / (rec["ms_per_step"] * 1e-3)) / rec["value"] < 1e-6
    rf = rec["roofline"]
    for k in ("bound", "kernel", "achieved", "peak", "unit", "frac", "traffic", "near_dominant"):
        assert k in rf, k
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and 0 < rf["frac"] < 1
    assert all(0 < s["GBps"] / rf["peak"] < 1.0 for s in rf["stages"].values() if s["GBps"])      # no stage above the measured peak
    cb = rec["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    e = rec["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 2 * 57_600_000 and e["d2h_bytes_per_step"] >= 2 * 57_600_000
    assert e["value"] < rec["value"]                                    # the host round trip cannot be faster than the resident step
    assert rec["gpu_launches"] >= 10 * rec["steps"]
    assert not set(rec["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
