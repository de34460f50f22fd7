"""`python -m tuun_b200 file.tuun`: the batch path of the reference's main.rs (--ui=false) — parse,
evaluate, optimize, play everything through the offline tracker, write float WAVs."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODULE = """
pi = 3.14159265;
$ = fn(freq_hz) => sine(2*pi * freq_hz, 0);
note = fn(f, dur) => $f | fin(time - dur);
// two programs; the second is quieter and captured
#{color=rgb(1,2,3)}
a = note(440, 0.25);
#{level_db=-6.0}
b = note(660, 0.5) | capture("b_note");
helper = 3;
"""


def run_cli(args, cwd):
    return subprocess.run([sys.executable, "-m", "tuun_b200"] + args, cwd=cwd, capture_output=True, text=True,
                          env=dict(os.environ, PYTHONPATH=ROOT), timeout=600)


def test_dry_run_lists_programs(tmp_path):
    src = tmp_path / "song.tuun"
    src.write_text(MODULE)
    r = run_cli([str(src), "--dry-run"], str(tmp_path))
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().split("\n")
    assert len(lines) == 2 and lines[0].startswith("a:") and lines[1].startswith("b:")
    bad = tmp_path / "bad.tuun"
    bad.write_text("#{color=rgb(0,0,0)}\nx = nope(1);")
    r = run_cli([str(bad), "--dry-run"], str(tmp_path))
    assert r.returncode == 1 and "Variable 'nope' not found" in r.stderr


@pytest.mark.gpu
def test_batch_render_writes_mix_and_captures(tmp_path):
    from oracle.binding import OracleProgram
    from tuun_b200.__main__ import programs_of
    from tuun_b200.frontend import Evaluator
    from tuun_b200.tracker import read_wav
    src = tmp_path / "song.tuun"
    src.write_text(MODULE)
    r = run_cli([str(src), "--output-dir", str(tmp_path / "out"), "--seconds", "2"], str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    mix, rate = read_wav(str(tmp_path / "out" / "mix.wav"))
    cap, _ = read_wav(str(tmp_path / "out" / "b_note.wav"))
    assert rate == 44100 and len(cap) == 22050
    progs = programs_of(str(src), Evaluator(44100, 90.0, None))
    want = np.zeros(len(mix), dtype=np.float32)
    for _, w in progs:
        row = OracleProgram(w, 44100).render(len(mix), block=1024)
        want[:len(row)] += row
    assert np.max(np.abs(mix - want)) <= 2e-4
    # -6 dB = 10^(-6/20) = 0.501: the captured note is the un-scaled one, the mix holds the scaled one
    assert abs(float(np.max(np.abs(cap))) - 1.0) < 1e-3
    assert 1.0 < float(np.max(np.abs(mix[:11025]))) <= 1.51
