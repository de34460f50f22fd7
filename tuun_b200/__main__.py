"""Batch rendering of a Tuun source file on the GPU — the `tuun <input_file> --ui=false` path of the
reference (src/main.rs:91-174): parse, evaluate, optimize and play every program of the file at
once through the (offline) tracker, then write what was mixed and what was captured as 32-bit
float WAV files.

    python -m tuun_b200 INPUT.tuun  [--library-root DIR] [--tempo 90] [--sample-rate 44100]
                                    [--seconds 10] [--output-dir DIR] [--dry-run]

INPUT.tuun is a module: every binding carrying at least one `#{...}` annotation is a program
(docs/language-spec.md, "Annotations"); `level_db=D` scales it like Player::play_program does
(player.rs:106-122, 265-288).  INPUT.tuunp (the older format of fm-variations.tuunp) holds one
program expression per line, evaluated after `open std`.
--dry-run stops after lowering (tb_lower_check): it needs no GPU.
"""
from __future__ import annotations

import argparse
import os
import re
import sys

import numpy as np

from .frontend import Env, Evaluator, evaluate, parse_module, parse_program, _bind, _slider_bindings
from .builder import to_waveform
from .optimizer import optimize
from .waveform import BinaryPointOp, Const, Marked, Operator, Waveform

MARK_TOP_LEVEL, MARK_AMPLITUDE, MARK_TERMINATOR = 1, 2, 3


def top_level(w: Waveform, level_db: float) -> Waveform:
    """build_top_level_waveform (player.rs:265-288): two constant multiplies under marks, added
    AFTER optimisation, so they are not folded."""
    amp = float(np.power(np.float32(10.0), np.float32(level_db) / np.float32(20.0), dtype=np.float32))
    return Marked(MARK_TOP_LEVEL, BinaryPointOp(
        Operator.Multiply, BinaryPointOp(Operator.Multiply, w, Marked(MARK_AMPLITUDE, Const(amp))),
        Marked(MARK_TERMINATOR, Const(1.0))))


def programs_of(path: str, ev: Evaluator):
    """[(display name, optimized + wrapped waveform)]."""
    text = open(path).read()
    out = []
    if path.endswith(".tuunp"):
        env, _ = ev.run_bindings(parse_module("open std;"))
        lines = [l for l in text.split("\n") if l.strip() and not l.strip().startswith("//")]
        for k, line in enumerate(lines):
            value = evaluate(parse_program(line), env)
            out.append((f"line {k + 1}", top_level(optimize(to_waveform(value)), 0.0)))
        return out
    env = ev.prelude
    for b in parse_module(text):
        if b[0] == "open":
            env, _ = ev.run_bindings([b], env)
            continue
        value = evaluate(b[2], env.extend(_slider_bindings(b[3])) if b[3] else env)
        frame = {}
        _bind(b[1], value, frame)
        env = env.extend(frame)
        if b[3]:  # annotated: a UI program
            level = 0.0
            for anno in b[3]:
                m = re.search(r"level_db\s*=\s*(-?[0-9.]+)", anno)
                if m:
                    level = float(m.group(1))
            try:
                w = to_waveform(value)
            except Exception:
                print(f"Program {len(out)} did not evaluate to a waveform", file=sys.stderr)
                continue
            name = b[1][1] if b[1][0] == "id" and b[1][1] != "_" else f"program {len(out)}"
            out.append((name, top_level(optimize(w), level)))
    return out


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m tuun_b200", description=__doc__.split("\n\n")[0])
    ap.add_argument("input_file")
    ap.add_argument("--library-root", default=None, help="directory holding std.tuun etc. (the reference's lib/v0)")
    ap.add_argument("--tempo", type=float, default=90.0)
    ap.add_argument("--sample-rate", type=int, default=44100)
    ap.add_argument("--buffer-size", type=int, default=1024)
    ap.add_argument("--seconds", type=float, default=10.0, help="stop after this long (infinite programs never end)")
    ap.add_argument("--output-dir", default=".")
    ap.add_argument("--device", type=int, default=-1)
    ap.add_argument("--dry-run", action="store_true", help="parse, evaluate, optimize and lower only (no GPU)")
    args = ap.parse_args(argv)
    root = args.library_root or os.path.join(os.path.dirname(os.path.abspath(args.input_file)), "lib", "v0")
    ev = Evaluator(args.sample_rate, args.tempo, root)
    from .frontend import EvalError, ParseError
    try:
        progs = programs_of(args.input_file, ev)
    except (EvalError, ParseError) as e:  # the reference prints the diagnostics and exits 1 (main.rs:80-89,141-150)
        print(f"{args.input_file}: Error: {e}", file=sys.stderr)
        return 1
    if not progs:
        print("no programs found", file=sys.stderr)
        return 1
    if args.dry_run:
        from .generator import lower_check
        for name, w in progs:
            info = lower_check(w)
            print(f"{name}: {info.n_nodes} nodes, {info.n_code_words // 4} code words, tile {info.tile}, "
                  f"{info.threads} threads/CTA, {info.smem_bytes} B shared")
        return 0
    from .tracker import OfflineTracker, write_wav
    t = OfflineTracker(args.sample_rate, args.buffer_size, device=args.device, max_seconds=args.seconds)
    for k, (name, w) in enumerate(progs):
        print(f"Playing program {name}")
        t.play(w, 0.0, id=k)
    mix = t.render_all(max_seconds=args.seconds)
    os.makedirs(args.output_dir, exist_ok=True)
    write_wav(os.path.join(args.output_dir, "mix.wav"), mix, args.sample_rate)
    for stem, samples in t.captured_output().items():
        write_wav(os.path.join(args.output_dir, f"{stem}.wav"), samples, args.sample_rate)
    print(f"All waveforms finished: {len(mix)} samples mixed, {len(t.captured)} captures, "
          f"{t.launches} kernel launches")
    return 0


if __name__ == "__main__":
    sys.exit(main())
