"""Host-side tests: the C ABI library loads and exports every symbol of include/mauve_b200.h, the C++ mirror of
the reference classes compiles, seed tables agree between C and Python.  No GPU, no compute calls."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()


def test_library_exports_every_declared_symbol():
    _build()
    import mauvealigner_b200 as mb
    hdr = open(os.path.join(ROOT, "include", "mauve_b200.h")).read()
    declared = set(re.findall(r"\b(mb_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"mb_params", "mb_result", "mb_stats", "mb_ctx", "mb_synth"}
    assert len(declared) >= 20
    lib = ctypes.CDLL(mb._lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert set(mb._lib.EXPORTS) <= declared


def test_no_cpu_fallback_without_gpu():
    import mauvealigner_b200 as mb
    if mb.lib().mb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(mb.MauveError):
        mb.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mauvealigner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.lower(), (f, "product code must not reference oracle/")
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert "oracle" not in open(p).read().lower()


def test_seed_tables_agree_c_and_python(tmp_path):
    import mauvealigner_b200 as mb
    src = tmp_path / "seeds.c"
    src.write_text('#include <stdio.h>\n#include "mauve_b200/seed_masks.h"\nint main(){int rk[5]={0,1,2,MB_CODING_SEED,MB_SOLID_SEED};'
                   'for(int w=3;w<=31;w++)for(int r=0;r<5;r++)printf("%d %d %llu %d\\n",w,r,(unsigned long long)mb_get_seed(w,rk[r]),'
                   'mb_seed_valid(mb_get_seed(w,rk[r])));for(unsigned long long l=1000;l<4000000000ULL;l*=3)printf("d %llu %d\\n",l,mb_default_seed_weight(l));return 0;}')
    exe = tmp_path / "seeds"
    subprocess.check_call(["gcc", "-O1", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split("\n")
    ranks = [0, 1, 2, mb.CODING_SEED, mb.SOLID_SEED]
    n = 0
    for line in out:
        t = line.split()
        if not t:
            continue
        if t[0] == "d":
            assert mb.default_seed_weight(int(t[1])) == int(t[2]), line
            continue
        w, r, pat, valid = int(t[0]), int(t[1]), int(t[2]), int(t[3])
        assert mb.get_seed(w, ranks[r]) == pat, line
        if pat:
            assert valid == 1 and mb.seed_valid(pat)
            assert mb.seed_weight(pat) == (w if w % 2 else w - 1)
        n += 1
    assert n == 29 * 5
    assert mb.default_seed_weight(5_000_000) == 15


def test_cpp_mirror_compiles(tmp_path):
    _build()
    exe = tmp_path / "compat_driver"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "compat_driver.cpp"), "-o", str(exe),
                           "-L", os.path.join(ROOT, "mauvealigner_b200"), "-lmauve_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "mauvealigner_b200")])
    assert exe.exists()


def test_python_mirror_shapes():
    import mauvealigner_b200 as mb
    m = mb.Match(3)
    m.SetLength(21)
    m.SetStart(0, 5)
    m.SetStart(2, -9)
    assert m.Multiplicity() == 2 and m.Orientation(0) == 0 and m.Orientation(2) == 1 and m[1] == mb.NO_MATCH
    assert m.Copy().Length() == 21 and m.LeftEnd(2) == 9
    ml = mb.MatchList()
    ml.seq_table = ["ACGT" * 100, "ACGA" * 100]
    ml.CreateMemorySMLs(0)
    assert mb.seed_valid(ml.seed_pattern) and ml.sml_table[0].SeedLength() == mb.seed_length(ml.seed_pattern)


def test_multiplicity_filter():
    import mauvealigner_b200 as mb
    ml = mb.MatchList()
    for starts in ([5, 0, -7], [1, 2, 3], [0, 0, 9], [4, -4, 0]):
        m = mb.Match(3)
        m.SetLength(21)
        for i, st in enumerate(starts):
            m.SetStart(i, st)
        ml.append(m)
    ml.MultiplicityFilter(2)
    assert [[m.Start(i) for i in range(3)] for m in ml] == [[5, 0, -7], [4, -4, 0]]
    ml.MultiplicityFilter(3)
    assert len(ml) == 0


def _random_match_list(rng, nseq, n, maxpos=400, maxlen=40):
    out = []
    for _ in range(n):
        ln = int(rng.integers(1, maxlen))
        st = []
        for g in range(nseq):
            r = rng.random()
            p = int(rng.integers(1, maxpos))
            st.append(0 if r < 0.25 else (p if r < 0.7 else -p))
        if sum(1 for x in st if x) >= 2:
            out.append((ln, st))
    return out


def _as_matchlist(mb, matches):
    ml = mb.MatchList()
    for ln, st in matches:
        m = mb.Match(len(st))
        m.SetLength(ln)
        for i, x in enumerate(st):
            m.SetStart(i, x)
        ml.append(m)
    return ml


def test_list_filters_python_cpp_and_per_base_restatement(tmp_path):
    """EliminateOverlaps and transposeMatches (src/mauveAligner.cpp:594-596,628-637; src/transposeCoordinates.cpp:46-65):
    the Python mirror, the C++ mirror (through the compiled driver, no GPU involved) and the per-base restatement of
    tests/brute.py agree on random match lists; hand-checked cases pin the rules themselves."""
    import numpy as np
    import mauvealigner_b200 as mb
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import brute
    _build()
    exe = tmp_path / "compat_driver"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "compat_driver.cpp"), "-o", str(exe),
                           "-L", os.path.join(ROOT, "mauvealigner_b200"), "-lmauve_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "mauvealigner_b200")])

    def write_list(path, matches, nseq):
        with open(path, "w") as f:
            f.write(f"FormatVersion\t3\nSequenceCount\t{nseq}\n")
            for i in range(nseq):
                f.write(f"Sequence{i}File\ts{i}\nSequence{i}Length\t1000\n")
            f.write(f"MatchCount\t{len(matches)}\n")
            for ln, st in matches:
                f.write("\t".join(str(x) for x in [ln] + st) + "\n")

    def cpp(args):
        out = subprocess.check_output([str(exe)] + args, text=True).strip()
        return [(int(r[0]), [int(x) for x in r[1:]]) for r in (l.split("\t") for l in out.split("\n") if l)]

    # hand-checked: [1..10] and [6..15] overlap in sequence 0 by 5 -> the second is cropped to [11..15]; its reverse
    # component in sequence 1 (left end 40, length 10) loses the RIGHT 5 bases, i.e. keeps left end 40
    ml = _as_matchlist(mb, [(10, [1, 20]), (10, [6, -40]), (4, [3, 70])])
    mb.EliminateOverlaps(ml)
    assert [(m.Length(), [m.Start(0), m.Start(1)]) for m in ml] == [(10, [1, 20]), (5, [11, -40])]
    # hand-checked: regions 101..110 and 201..300 (filtered 1..10, 11..110); a forward match at filtered 8, length 6
    # splits into original 108..110 and 201..203; sequence 1 follows column by column
    ml = _as_matchlist(mb, [(6, [8, 50])])
    mb.transposeMatches(ml, 0, [101, 110, 201, 300])
    assert [(m.Length(), [m.Start(0), m.Start(1)]) for m in ml] == [(3, [108, 50]), (3, [201, 53])]
    ml = _as_matchlist(mb, [(6, [-8, 50])])
    mb.transposeMatches(ml, 0, [101, 110, 201, 300])
    assert [(m.Length(), [m.Start(0), m.Start(1)]) for m in ml] == [(3, [-108, 53]), (3, [-201, 50])]

    rng = np.random.default_rng(5)
    for trial in range(12):
        nseq = 2 + trial % 3
        matches = _random_match_list(rng, nseq, 30 + 10 * trial)
        lst = tmp_path / f"l{trial}.mums"
        write_list(lst, matches, nseq)
        ml = _as_matchlist(mb, matches)
        mb.EliminateOverlaps(ml)
        got = [(m.Length(), [m.Start(g) for g in range(nseq)]) for m in ml]
        assert got == brute.eliminate_overlaps(matches)
        assert got == cpp(["overlaps", str(lst)])
        for g in range(nseq):  # no two matches overlap in any sequence any more
            iv = sorted((abs(st[g]), abs(st[g]) + ln - 1) for ln, st in got if st[g])
            assert all(a[1] < b[0] for a, b in zip(iv, iv[1:]))
        regions = []
        x = 1
        for _ in range(6):
            x += int(rng.integers(0, 50))
            w = int(rng.integers(20, 120))
            regions += [x, x + w - 1]
            x += w
        seqI = trial % nseq
        ml = _as_matchlist(mb, matches)
        mb.transposeMatches(ml, seqI, regions)
        got = [(m.Length(), [m.Start(g) for g in range(nseq)]) for m in ml]
        assert got == brute.transpose_matches(matches, seqI, regions)
        assert got == cpp(["transpose", str(lst), str(seqI)] + [str(r) for r in regions])
        assert sum(ln for ln, _ in got) == sum(ln for ln, _ in matches)  # splitting only


def test_multi_gpu_entry_points_reject_bad_arguments():
    """argument checks of the multi-GPU entry points run before any device work (no GPU needed)"""
    import ctypes as C
    import mauvealigner_b200 as mb
    from mauvealigner_b200 import _lib as L
    lib = mb.lib()
    p = L.MbParams(L.MODE_UNIQUE, 0, 2, 1000, 0)
    assert lib.mb_find_multi(None, 2, C.byref(p)) == -1            # MB_E_ARG: no contexts
    arr = (C.c_void_p * 2)(None, None)
    assert lib.mb_find_multi(arr, 2, C.byref(p)) == -1             # null context
    assert lib.mb_find_multi(arr, 0, C.byref(p)) == -1             # world < 1
    assert lib.mb_dist_push(None, None, None, 1, None, None) == -1
    assert lib.mb_dist_rows_pack(None, None, None, None) == -1
    assert lib.mb_dist_match_pack(None, None, None, None, None, None, None) == -1


def test_product_library_holds_no_generator_and_reference_arm_maps_no_product_library():
    """The synthetic-genome generator lives in tools/synth/libmbsynth.so, and `bench.py --impl reference` (the CPU arm)
    must run without mapping libmauve_b200.so."""
    _build()
    import mauvealigner_b200 as mb
    lib = ctypes.CDLL(mb._lib.LIB_PATH)
    assert not hasattr(lib, "mb_synth_create")
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--config', '1', '--ref-scale', '200', '--steps', '1', '--warmup', '0'];"
            "runpy.run_path('bench.py', run_name='__main__');"
            "maps = open('/proc/self/maps').read();"
            "assert 'libmauve_b200' not in maps, 'product library mapped by the reference arm';"
            "assert 'liboracle' in maps and 'libmbsynth' in maps")
    import subprocess
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.config_dict(1, 1)  # the same `config` the repo arm prints
