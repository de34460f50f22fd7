"""The front end (tuun_b200/frontend.py): the reference evaluator's own known-answer tests
(src/lib/eval.rs:515-688) restated, the precedence rules of its parser, and — when the reference
tree is available — its real `lib/v0/std.tuun` and `fm-variations.tuunp` sources evaluated to the
very trees the hand-written workloads build."""
import os
import re

import numpy as np
import pytest

from tuun_b200 import workloads as W
from tuun_b200.frontend import EvalError, Evaluator, ParseError, evaluate, parse_module, parse_program, show
from tuun_b200.waveform import Append, BinaryPointOp, Const, Fin, Fixed, Operator, Sine, Time, add, mul, sub

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "lib", "v0")),
                               reason="reads the reference's Tuun library sources (not present on the GPU box)")


def run(src, **kw):
    return Evaluator(**kw).evaluate_source(src)


def err(src, **kw):
    with pytest.raises(EvalError) as e:
        run(src, **kw)
    return str(e.value)


def test_named_arguments():  # eval.rs:515-583
    f = "let f = fn(x, y = 10) => x * y + 1 in "
    assert show(run(f + "f(2)")) == "21" and show(run(f + "f(2, y = 5)")) == "11"
    assert err(f + "f(2, 3)") == "extra positional parameter"
    assert err(f + "f(2, z = 3)") == 'no named parameter "z"'
    assert err(f + "f(y = 2)") == 'missing parameter "x"'
    g = "let g = fn(y = 1) => y in "
    assert show(run(g + "g()")) == "1" and show(run(g + "g(y = 3)")) == "3"
    assert show(run("let a = 5, f = fn(x, y = a * 2) => x + y in f(1)")) == "11"
    assert show(run("let y = 100, f = fn(x, y = 10) => x * y in f(2)")) == "20"
    h = "let f = fn((a, b), y = 1) => a + b + y in "
    assert show(run(h + "f((1, 2))")) == "4" and show(run(h + "f((1, 2), y = 10)")) == "13"
    assert 'built-in "sine"' in err("sine(440, y = 1)")


def test_named_defaults_evaluate_once():  # eval.rs:585-621
    printed = []
    ev = Evaluator(print_fn=printed.append)
    assert show(ev.evaluate_source("let f = fn(x, y = debug(1)) => x, _ = f(1), _ = f(2) in f(3)")) == "3"
    assert printed == ["debug: [1]"]
    printed.clear()
    ev.evaluate_source("let f = fn(x, y = debug(1)) => x in 0")
    assert printed == ["debug: [1]"]


def test_opens_are_scoped():  # eval.rs:623-651
    ev = Evaluator(modules={"b": "two = 2;", "a": "open b; alias = two;"})
    assert show(ev.evaluate_source("alias", "open a;")) == "2"
    with pytest.raises(EvalError) as e:
        ev.evaluate_source("two", "open a;")
    assert str(e.value) == "Variable 'two' not found in context"


def test_application_arity_and_closures():  # eval.rs:653-687
    assert err("(fn(x) => x)(2, 3)") == "extra positional parameter"
    assert err("(fn(x, y) => x)(2)") == 'missing parameter "y"'
    assert show(run("(fn((y, z)) => (z, y))((4, 5))")) == "(5, 4)"
    assert err("(fn((y, z)) => y)(4, 5)") == "extra positional parameter"
    assert show(run("(fn(x) => fn(x) => x)(7)(5)")) == "5"
    assert show(run("(fn(x) => fn(y, z) => (x, y, z))(3)(4, 5)")) == "(3, 4, 5)"
    assert show(run("(fn(x, (y, z)) => (x, y, z))(3, (4, 5))")) == "(3, 4, 5)"


def test_precedence_and_sugar():  # parser.rs:641-815
    assert show(run("1 + 2 * 3")) == "7" and show(run("(1 + 2) * 3")) == "9"
    assert show(run("10 - 4 - 3")) == "3" and show(run("-2 * 3")) == "-6" and show(run("2 * -3")) == "-6"
    assert run("1 + 2 == 3") is True and run("2 * 2 < 3") is False
    assert show(run("3 | fn(x) => x + 1")) == "4"                      # reverse application
    assert show(run("1 + 2 | fn(x) => x * 2")) == "6"                  # `|` is looser than `+`
    assert show(run("if 1 < 2 then 10 else 20")) == "10"
    assert show(run("let (a, b) = (1, 2), in a + b")) == "3"           # tuple pattern, trailing comma
    assert show(run("map(fn(x) => x * x, [1, 2, 3])")) == "[1, 4, 9]"
    assert show(run("reduce(fn(acc, x) => acc + x, 0, unfold(fn(i) => i + 1, 1, 4))")) == "10"
    assert show(run("nth(1, append([1], [2, 3]))")) == "2"
    assert run("2 * time") == mul(Const(2.0), Time())                   # float * waveform promotes
    assert run("time - 0.5 | fin | fn(f) => f(1)") == Fin(sub(Time(), Const(0.5)), Const(1.0))
    assert run("fixed([1, 2]) // trailing comment") == Fixed([1, 2])
    assert run("{[1, time]}") == BinaryPointOp(Operator.Merge, Const(1.0),
                                               BinaryPointOp(Operator.Merge, Time(), Fin(Const(0.0), Const(0.0))))
    seq2 = run("<[1 | fin(time - 1) | seq(time - 1), 2 | fin(time - 1)]>")
    assert seq2 == BinaryPointOp(Operator.Merge, Fin(sub(Time(), Const(1.0)), Const(1.0)),
                                 Append(Fin(sub(Time(), Const(1.0)), Const(0.0)), Fin(sub(Time(), Const(1.0)), Const(2.0))))
    with pytest.raises(ParseError):
        parse_program("1 +")
    with pytest.raises(ParseError):
        parse_program("fn(x, y = 1, z) => x")  # positional after named (parser.rs:303-313)
    mod = parse_module('#{level_db=-3.0} a = 1; open std; (b, c) = (2, 3);')
    assert [b[0] for b in mod] == ["def", "open", "def"] and mod[0][3] == ["level_db=-3.0"]


@needs_ref
def test_std_sources_evaluate_to_the_workload_trees():
    ev = Evaluator(44100, 120, os.path.join(REF, "lib", "v0"))
    assert ev.waveform("$440 * Qw", "open std;") == W.cfg1_from_source()
    assert ev.waveform("let h = harmonica(Q, 440) in <[h, h, h, h]>", "open std;") == W.cfg2_harmonica(4)
    lines = [l for l in open(os.path.join(REF, "fm-variations.tuunp")).read().split("\n")
             if l.strip() and not l.strip().startswith("//")]
    assert len(lines) == 12
    for (name, want), line in zip(W.cfg3_fm_variations(), lines):
        src = re.sub(r'\|\s*capture\("[^"]*"\)', "", line)
        assert ev.waveform(src, "open std;") == want, name
    # the un-optimized bench program of benches/tracker_benches.rs:141 (its own mini-library)
    ctx = """
    pi = 3.14159265;
    $ = fn(freq_hz) => sine(2*pi * freq_hz, 0);
    triangle = fn(freq_hz) => let t = $freq_hz, slope = 4 * freq_hz, a = time * slope - 1, b = time * -slope + 3 in alt(t, reset(t, a), reset(t, b));
    linear = fn(initial, slope) => initial + (time * slope);
    Rw = fn(dur, level) => linear(level, -level / dur) | fin(time - dur);
    R = fn(dur, level) => fn(w) => w * Rw(dur, level);"""
    large = Evaluator().waveform("triangle(55) + (noise * 0.2) | R(1.0, 1.0)", ctx, optimize=False)
    assert large == dict((n, w) for n, w, _ in W.tracker_benches())["large_440"]


@needs_ref
def test_every_library_module_parses_and_evaluates():
    root = os.path.join(REF, "lib", "v0")
    ev = Evaluator(44100, 90, root)
    for f in sorted(os.listdir(root)):
        if f.endswith(".tuun"):
            mod = parse_module(open(os.path.join(root, f)).read())
            assert len(mod) >= 3, f
            ev.run_bindings(mod)  # every definition evaluates
    # captures survive as Captured nodes
    w = ev.waveform('$440 | fin(time - 1) | capture("a")', "open std;", optimize=False)
    assert type(w).__name__ == "Captured" and w.file_stem == "a"
