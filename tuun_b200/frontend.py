"""Front end: Tuun source text -> `Waveform` trees for the renderer.

A from-scratch recursive-descent parser and environment-passing evaluator for the Tuun
expression language (docs/language-spec.md), following the grammar of the reference's nom parser
(src/lib/parser.rs:98-830) and the call-by-value semantics of its evaluator (src/lib/eval.rs:
named parameters with defaults evaluated once, tuple patterns, `open` with non-transitive
exports, built-ins that reject named arguments).  The values the built-ins produce are those of
tuun_b200.builder (which mirrors builtins.rs); modules such as `std` are read as Tuun SOURCE from a
library root at run time — nothing of lib/v0 is copied into this repository.

Precedence, loosest first (parser.rs:641-815):   `\\`   `|`   == != <= >= < >   + - &   * /
application `f(x, name = y)`   unary `- $ @ ! % ?` on a primitive.  `{e}` is `__chord(e)`, `<e>` is
`__sequence(e)`, `let p = e, ... in body` is nested application, `-1` is `-`(1).
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from . import builder as B
from .waveform import Fixed, Noise, Time, Waveform

F = np.float32
KEYWORDS = {"fn", "let", "in", "if", "then", "else", "open"}
UNARY = "!@$%-?"


class ParseError(Exception):
    def __init__(self, message: str, pos: int, text: str):
        line = text.count("\n", 0, pos) + 1
        col = pos - (text.rfind("\n", 0, pos) + 1) + 1
        super().__init__(f"{line}:{col}: {message}")
        self.pos = pos


class EvalError(Exception):
    pass


# ---------------------------------------------------------------------------------------------
# AST: ("num", f32) ("str", s) ("var", name) ("fn", [pattern], [(name, expr)], body)
#      ("app", fn, [expr], [(name, expr)]) ("if", c, t, e) ("tuple", [expr]) ("list", [expr])
# patterns: ("id", name) | ("ptuple", [pattern])
# ---------------------------------------------------------------------------------------------
_FLOAT = re.compile(r"(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)")
_IDENT = re.compile(r"(_?[A-Za-z0-9]+[A-Za-z0-9_#]*)")
_DUNDER = re.compile(r"__[A-Za-z0-9_#]*")
_TRIVIA = re.compile(r"(\s+|//[^\n]*)*")


class Parser:
    def __init__(self, text: str):
        self.t = text
        self.i = 0

    # -- lexical helpers
    def ws(self):
        self.i = _TRIVIA.match(self.t, self.i).end()

    def peek(self, s: str) -> bool:
        return self.t.startswith(s, self.i)

    def eat(self, s: str) -> bool:
        if self.peek(s):
            self.i += len(s)
            return True
        return False

    def expect(self, s: str, what: str):
        if not self.eat(s):
            raise ParseError(what, self.i, self.t)

    def fail(self, msg: str):
        frag = self.t[self.i:self.i + 30].split("\n")[0]
        raise ParseError(f"{msg} (at '{frag}')" if frag else f"{msg} (at end of input)", self.i, self.t)

    def keyword(self, kw: str) -> bool:
        m = _IDENT.match(self.t, self.i)
        if m and m.group(0) == kw:
            self.i = m.end()
            return True
        return False

    def identifier(self) -> Optional[str]:
        """parser.rs:179-201: a name, a unary operator, or a lone underscore."""
        m = _IDENT.match(self.t, self.i)
        if m and m.group(0) not in KEYWORDS and not m.group(0).startswith("__"):
            self.i = m.end()
            return m.group(0)
        if self.i < len(self.t) and self.t[self.i] in UNARY:
            self.i += 1
            return self.t[self.i - 1]
        if self.peek("_") and not re.match(r"_[A-Za-z0-9_]", self.t[self.i:self.i + 2]):
            self.i += 1
            return "_"
        return None

    # -- patterns and parameters
    def pattern(self):
        if self.eat("("):
            self.ws()
            items = []
            if not self.peek(")"):
                while True:
                    items.append(self.pattern())
                    self.ws()
                    if not self.eat(","):
                        break
                    self.ws()
            self.expect(")", "expected ')' at end of tuple pattern")
            return ("ptuple", items)
        name = self.identifier()
        if name is None:
            self.fail("expected a pattern")
        return ("id", name)

    def named_item(self):
        """`name = expr` (not `name == ...`); returns None without consuming if it is not one."""
        save = self.i
        name = self.identifier()
        if name is not None:
            self.ws()
            if self.peek("=") and not self.peek("=="):
                self.i += 1
                self.ws()
                return name, self.expr()
        self.i = save
        return None

    # -- primitives (parser.rs:514-537)
    def primitive(self):
        t = self.t
        m = _FLOAT.match(t, self.i)  # tried first, like parse_literal (a leading '-' is the unary operator)
        if m:
            self.i = m.end()
            return ("num", F(m.group(0)))
        if self.peek('"'):
            end = t.find('"', self.i + 1)
            if end < 0:
                self.fail("unterminated string")
            s = t[self.i + 1:end]
            self.i = end + 1
            return ("str", s)
        save = self.i
        if self.keyword("fn"):
            self.ws()
            if self.peek("("):
                return self.function()
            self.i = save
        if self.keyword("let"):
            return self.let()
        if self.keyword("if"):
            return self.if_then_else()
        if self.i < len(t) and t[self.i] in UNARY:  # unary application binds to a primitive
            op = t[self.i]
            self.i += 1
            return ("app", ("var", op), [self.primitive()], [])
        m = _DUNDER.match(t, self.i)
        if m:
            self.i = m.end()
            return ("var", m.group(0))
        name = self.identifier()
        if name is not None:
            if name == "_":
                self.i -= 1
                self.fail("'_' may be bound but not referenced")
            return ("var", name)
        if self.eat("{"):
            self.ws()
            e = self.expr()
            self.ws()
            self.expect("}", "expected '}' at end of chord")
            return ("app", ("var", "__chord"), [e], [])
        if self.eat("<"):
            self.ws()
            e = self.expr()
            self.ws()
            self.expect(">", "expected '>' at end of sequence")
            return ("app", ("var", "__sequence"), [e], [])
        if self.eat("("):
            items = self.expr_list(")", "expected ')' at end of tuple")
            return items[0] if len(items) == 1 else ("tuple", items)
        if self.eat("["):
            return ("list", self.expr_list("]", "expected ']' at end of list"))
        self.fail("unexpected input")

    def expr_list(self, close: str, what: str):
        self.ws()
        items = []
        if not self.peek(close):
            while True:
                items.append(self.expr())
                self.ws()
                if not self.eat(","):
                    break
                self.ws()
        self.expect(close, what)
        return items

    def function(self):
        self.expect("(", "expected '(' after fn")
        self.ws()
        positional, named, names = [], [], []
        if not self.peek(")"):
            while True:
                item = self.named_item()
                if item is not None:
                    if item[0] in names:
                        raise ParseError(f'named parameter "{item[0]}" appears more than once', self.i, self.t)
                    names.append(item[0])
                    named.append(item)
                else:
                    if named:
                        raise ParseError("positional arguments should appear before named ones", self.i, self.t)
                    p = self.pattern()
                    _pattern_names(p, names)
                    positional.append(p)
                self.ws()
                if not self.eat(","):
                    break
                self.ws()
        self.expect(")", "expected ')' at end of parameter list")
        self.ws()
        self.expect("=>", "expected '=>'")
        self.ws()
        return ("fn", positional, named, self.expr())

    def let(self):
        bindings = []
        while True:
            self.ws()
            self.annotations()
            if self.keyword("in"):  # optional trailing comma
                break
            p = self.pattern()
            self.ws()
            self.expect("=", "expected '=' in definition")
            self.ws()
            bindings.append((p, self.expr()))
            self.ws()
            if self.eat(","):
                continue
            if not self.keyword("in"):
                self.fail("expected 'in'")
            break
        self.ws()
        body = self.expr()
        for p, e in reversed(bindings):  # de-sugars to nested applications (parser.rs:445-452)
            body = ("app", ("fn", [p], [], body), [e], [])
        return body

    def if_then_else(self):
        self.ws()
        c = self.expr()
        self.ws()
        if not self.keyword("then"):
            self.fail("expected 'then'")
        self.ws()
        a = self.expr()
        self.ws()
        if not self.keyword("else"):
            self.fail("expected 'else'")
        self.ws()
        return ("if", c, a, self.expr())

    # -- application and operators
    def application(self):
        e = self.primitive()
        while True:
            save = self.i
            self.ws()
            if not self.eat("("):
                self.i = save
                return e
            self.ws()
            positional, named = [], []
            if not self.peek(")"):
                while True:
                    item = self.named_item()
                    if item is not None:
                        if any(n == item[0] for n, _ in named):
                            raise ParseError(f'named parameter "{item[0]}" appears more than once', self.i, self.t)
                        named.append(item)
                    else:
                        if named:
                            raise ParseError("positional arguments should appear before named ones", self.i, self.t)
                        positional.append(self.expr())
                    self.ws()
                    if not self.eat(","):
                        break
                    self.ws()
            self.expect(")", "expected ')' at end of arguments")
            e = ("app", e, positional, named)

    def _fold(self, sub, ops):
        e = sub()
        while True:
            save = self.i
            self.ws()
            for op in ops:
                if self.peek(op):
                    self.i += len(op)
                    self.ws()
                    rhs = sub()
                    e = ("app", ("var", op), [e, rhs], [])
                    break
            else:
                self.i = save
                return e

    def multiplicative(self):
        return self._fold(self.application, ("*", "/"))

    def additive(self):
        return self._fold(self.multiplicative, ("+", "-", "&"))

    def relational(self):
        e = self.additive()
        while True:
            save = self.i
            self.ws()
            for op in ("==", "!=", "<=", ">=", "<", ">"):
                if self.peek(op):
                    mark = self.i
                    self.i += len(op)
                    self.ws()
                    try:
                        rhs = self.additive()
                    except ParseError:
                        self.i = mark  # e.g. the '>' that closes a sequence
                        break
                    e = ("app", ("var", op), [e, rhs], [])
                    save = None
                    break
            if save is not None:
                self.i = save
                return e

    def reverse_application(self):
        arg = self.relational()
        while True:
            save = self.i
            self.ws()
            if self.eat("|"):
                self.ws()
                f = self.relational()
                arg = ("app", f, [arg], [])
            else:
                self.i = save
                return arg

    def expr(self):
        e = self.reverse_application()
        while True:
            save = self.i
            self.ws()
            if self.eat("\\"):
                self.ws()
                rhs = self.reverse_application()
                e = ("app", ("var", "\\"), [e, rhs], [])
            else:
                self.i = save
                return e

    # -- modules
    def annotations(self) -> List[str]:
        """`#{...}` sets before a binding (parser.rs:1073-1111); kept as raw text."""
        out = []
        while self.peek("#{"):
            depth, j = 0, self.i + 1
            while j < len(self.t):
                if self.t[j] in "{[(":
                    depth += 1
                elif self.t[j] in "}])":
                    depth -= 1
                    if depth == 0:
                        break
                j += 1
            if j >= len(self.t):
                self.fail("unterminated annotation")
            out.append(self.t[self.i + 2:j])
            self.i = j + 1
            self.ws()
        return out

    def module(self):
        """binding ';' ...  ->  [("open", [path]) | ("def", pattern, expr, [annotations])]"""
        out = []
        while True:
            self.ws()
            if self.i >= len(self.t):
                return out
            if self.eat(";"):
                continue
            annos = self.annotations()
            save = self.i
            if self.keyword("open"):
                self.ws()
                path = [self.identifier()]
                while self.eat("."):
                    path.append(self.identifier())
                if None in path:
                    self.fail("expected a module path after 'open'")
                out.append(("open", path))
            else:
                self.i = save
                p = self.pattern()
                self.ws()
                self.expect("=", "expected '=' in definition")
                self.ws()
                out.append(("def", p, self.expr(), annos))
            self.ws()
            if self.i < len(self.t) and not self.eat(";"):
                self.fail("expected ';' after binding")


def _pattern_names(p, names: List[str]):
    if p[0] == "id":
        if p[1] != "_" and p[1] in names:
            raise EvalError(f'parameter "{p[1]}" appears more than once')
        names.append(p[1])
    else:
        for q in p[1]:
            _pattern_names(q, names)


def parse_program(text: str):
    """One expression (parser.rs:848-877)."""
    p = Parser(text)
    p.ws()
    e = p.expr()
    p.ws()
    if p.i != len(text):
        p.fail("unexpected input")
    return e


def parse_module(text: str):
    return Parser(text).module()


# ---------------------------------------------------------------------------------------------
# Evaluation (eval.rs): values are np.float32, str, bool, tuple, list, Waveform, builder.Seq,
# Closure and BuiltIn.
# ---------------------------------------------------------------------------------------------
@dataclass
class Closure:
    positional: list
    named: List[Tuple[str, object]]  # defaults, already evaluated (eval.rs:236-247)
    body: object
    env: "Env"


@dataclass
class BuiltIn:
    name: str
    fn: Callable


class Env:
    """An immutable chain of frames; later bindings shadow earlier ones."""

    def __init__(self, names: Optional[Dict[str, object]] = None, parent: Optional["Env"] = None):
        self.names = names or {}
        self.parent = parent

    def lookup(self, name: str):
        e = self
        while e is not None:
            if name in e.names:
                return e.names[name]
            e = e.parent
        raise EvalError(f"Variable '{name}' not found in context")

    def extend(self, names: Dict[str, object]) -> "Env":
        return Env(names, self)


def _bind(pattern, value, into: Dict[str, object]):
    """extend_context (eval.rs:150-199)."""
    if pattern[0] == "id":
        into[pattern[1]] = value
        return
    if not isinstance(value, tuple):
        raise EvalError(f"Pattern does not match actual expression {show(value)}")
    if len(pattern[1]) != len(value):
        raise EvalError("Mismatched number of elements in pattern and arguments")
    for p, v in zip(pattern[1], value):
        _bind(p, v, into)


def _show_pattern(p) -> str:
    return p[1] if p[0] == "id" else "(" + ", ".join(_show_pattern(q) for q in p[1]) + ")"


def apply(f, positional: list, named: Optional[List[Tuple[str, object]]] = None):
    named = named or []
    if isinstance(f, Closure):
        for k, (name, _) in enumerate(named):
            if any(n == name for n, _ in named[:k]):
                raise EvalError(f'named parameter "{name}" appears more than once')
            if not any(n == name for n, _ in f.named):
                raise EvalError(f'no named parameter "{name}"')
        if len(positional) > len(f.positional):
            raise EvalError("extra positional parameter")
        if len(positional) < len(f.positional):
            raise EvalError(f'missing parameter "{_show_pattern(f.positional[len(positional)])}"')
        frame: Dict[str, object] = {}
        for p, v in zip(f.positional, positional):
            _bind(p, v, frame)
        given = dict(named)
        for name, default in f.named:
            frame[name] = given.get(name, default)
        return evaluate(f.body, f.env.extend(frame))
    if isinstance(f, BuiltIn):
        if named:
            raise EvalError(f'named argument "{named[0][0]}" is not supported by built-in "{f.name}"')
        try:
            return f.fn(*positional)
        except B.TuunError as e:
            raise EvalError(str(e)) from None
        except TypeError as e:
            raise EvalError(f"Invalid arguments for {f.name}: {e}") from None
    raise EvalError(f"Invalid application: {show(f)}")


def evaluate(e, env: Env):
    k = e[0]
    if k == "num":
        return e[1]
    if k == "str":
        return e[1]
    if k == "var":
        return env.lookup(e[1])
    if k == "fn":
        return Closure(e[1], [(n, evaluate(d, env)) for n, d in e[2]], e[3], env)
    if k == "app":
        f = evaluate(e[1], env)
        pos = [evaluate(a, env) for a in e[2]]
        named = [(n, evaluate(a, env)) for n, a in e[3]]
        return apply(f, pos, named)
    if k == "if":
        c = evaluate(e[1], env)
        if c is True:
            return evaluate(e[2], env)
        if c is False:
            return evaluate(e[3], env)
        raise EvalError("Expected boolean condition")
    if k == "tuple":
        return tuple(evaluate(a, env) for a in e[1])
    if k == "list":
        return [evaluate(a, env) for a in e[1]]
    raise EvalError(f"unknown expression kind {k}")


def show(v) -> str:
    if isinstance(v, (np.floating, float)):
        x = float(v)
        return str(int(x)) if x == int(x) and abs(x) < 1e15 else repr(float(F(x)))
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, list):
        return "[" + ", ".join(show(x) for x in v) + "]"
    if isinstance(v, tuple):
        return "(" + ", ".join(show(x) for x in v) + ")"
    if isinstance(v, Closure):
        return "fn(...)"
    if isinstance(v, BuiltIn):
        return v.name
    return str(v)


# ---------------------------------------------------------------------------------------------
# Built-ins (builtins.rs:1010-1066) over the value layer of tuun_b200.builder
# ---------------------------------------------------------------------------------------------
def _num(x) -> bool:
    return isinstance(x, (np.floating, float, int)) and not isinstance(x, bool)


def _cmp(name, op):
    def f(a, b):
        if _num(a) and _num(b):
            return bool(op(F(a), F(b)))
        raise B.TuunError(f"Invalid arguments for {name}")
    return f


def _equals(neg):
    def f(a, b):
        same = (isinstance(a, bool) and isinstance(b, bool)) or (_num(a) and _num(b)) or \
               (isinstance(a, str) and isinstance(b, str))
        if not same:
            raise B.TuunError("Invalid arguments for " + ("!=" if neg else "=="))
        return bool((a != b) if neg else (a == b))
    return f


def _map(f, xs):
    if not isinstance(xs, list):
        raise B.TuunError("Invalid arguments for map")
    return [apply(f, [x]) for x in xs]


def _reduce(f, acc, xs):
    if not isinstance(xs, list):
        raise B.TuunError("Invalid arguments for reduce")
    for x in xs:
        acc = apply(f, [acc, x])
    return acc


def _unfold(f, seed, n):
    if not (_num(n) and float(n) >= 0 and float(n) == int(n)):
        raise B.TuunError("Invalid arguments for unfold")
    out, cur = [], seed
    for _ in range(int(n)):
        out.append(cur)
        cur = apply(f, [cur])
    return out


def _append(first, *rest):
    from .waveform import Append
    if isinstance(first, list):
        out = list(first)
        for r in rest:
            if not isinstance(r, list):
                raise B.TuunError("Expected more lists as arguments for append")
            out.extend(r)
        return out
    if isinstance(first, Waveform):
        out = first
        for r in rest:
            if not isinstance(r, Waveform):
                raise B.TuunError("Expected more waveforms as arguments for append")
            out = Append(out, r)
        return out
    raise B.TuunError("Invalid arguments for append")


def _nth(i, xs):
    if not (_num(i) and isinstance(xs, list)):
        raise B.TuunError("Invalid arguments for nth")
    k = int(float(i)) if float(i) >= 0 else 0  # Rust `as usize` saturates
    if k >= len(xs):
        raise B.TuunError(f"No element with index {show(i)}")
    return xs[k]


def _fixed(xs):
    if not (isinstance(xs, list) and all(_num(x) for x in xs)):
        raise B.TuunError("Invalid argument for fixed waveform")
    return Fixed([float(F(x)) for x in xs])


def _log(value, base):
    if _num(value) and _num(base):
        with np.errstate(all="ignore"):
            return F(np.log(F(value), dtype=F) / np.log(F(base), dtype=F))
    raise B.TuunError("Invalid arguments for log")


def _sqrt(x):
    if _num(x) and float(x) >= 0:
        return F(np.sqrt(F(x), dtype=F))
    raise B.TuunError("Invalid argument for sqrt")


def _exp(x):
    if _num(x):
        return F(np.exp(F(x), dtype=F))
    raise B.TuunError("Invalid argument for exp")


def _minus(*args):
    if len(args) == 1:
        return B.minus(args[0])
    return B.minus(*args)


def builtin_bindings(print_fn: Callable[[str], None] = print) -> Dict[str, object]:
    def debug(*args):
        print_fn("debug: [" + ", ".join(show(a) for a in args) + "]")
        return args[-1] if args else []

    table = {
        "+": B.plus, "-": _minus, "*": B.times, "/": B.divide, "&": B.merge, "\\": B.followed_by,
        "==": _equals(False), "!=": _equals(True),
        "<": _cmp("<", lambda a, b: a < b), "<=": _cmp("<=", lambda a, b: a <= b),
        ">": _cmp(">", lambda a, b: a > b), ">=": _cmp(">=", lambda a, b: a >= b),
        "pow": B.power, "log": _log, "sqrt": _sqrt, "exp": _exp, "sine": B.sine, "cos": B.cos,
        "map": _map, "reduce": _reduce, "unfold": _unfold, "append": _append, "nth": _nth, "fixed": _fixed,
        "fin": B.fin, "seq": B.seq, "unseq": B.unseq, "filter": B.filter_, "reset": B.reset, "alt": B.alt,
        "capture": B.capture, "mark": B.mark, "__chord": B.chord, "__sequence": B.sequence, "debug": debug,
    }
    names: Dict[str, object] = {k: BuiltIn(k, v) for k, v in table.items()}
    # curried built-ins return Python callables: wrap them so they can be applied
    for k in ("fin", "seq", "filter", "capture", "mark", "unseq"):
        inner = table[k]
        names[k] = BuiltIn(k, (lambda inner, k: lambda *a: BuiltIn(f"{k}(..)", inner(*a)))(inner, k))
    names.update({"true": True, "false": False, "time": Time(), "noise": Noise()})
    return names


def _slider_bindings(annotations: List[str]) -> Dict[str, object]:
    """`sliders=["label:normalized:min:max", ...]` (parser.rs:938-1038, slider.rs:56-82): each label is
    bound, for that one program, to Marked(label, Const(min + normalized * (max - min)))."""
    from .waveform import Const, Marked
    out: Dict[str, object] = {}
    for anno in annotations:
        m = re.search(r"sliders\s*=\s*\[(.*?)\]", anno, re.S)
        if not m:
            continue
        for k, entry in enumerate(re.findall(r'"([^"]*)"', m.group(1))):
            parts = entry.split(":", 2)
            if len(parts) < 3:
                continue
            try:
                norm = F(parts[1])
                if parts[2].lstrip().startswith("fn"):  # user-defined: value = (fn)(normalized), slider.rs:28-52
                    value = evaluate(parse_program(f"({parts[2]})({float(norm)!r})"), Env(builtin_bindings()))
                else:
                    lo, hi = (F(x) for x in parts[2].split(":")[:2])
                    value = lo + norm * (hi - lo)
            except (ValueError, IndexError):
                continue
            out[parts[0]] = Marked(1000 + k, Const(float(F(value))))
    return out


class Evaluator:
    """`Evaluator::new(sample_rate, tempo, library_root)` (src/lib/evaluator.rs:107-137): the
    prelude binds the built-ins plus `tempo` and `sample_rate`; `open a.b` reads
    `<library_root>/a/b.tuun`; every module (and program) starts with an implicit `open __prelude`."""

    def __init__(self, sample_rate: int = 44100, tempo: float = 120.0, library_root: Optional[str] = None,
                 print_fn: Callable[[str], None] = print, modules: Optional[Dict[str, str]] = None):
        self.library_root = library_root
        self.module_sources = dict(modules or {})  # dotted path -> source text (embedded modules)
        names = builtin_bindings(print_fn)
        names["tempo"] = F(tempo)
        names["sample_rate"] = F(sample_rate)
        self.prelude = Env(names)
        self._modules: Dict[str, Dict[str, object]] = {}

    def _module_exports(self, path: List[str]) -> Dict[str, object]:
        key = ".".join(path)
        if key not in self._modules:
            if key in self.module_sources:
                text = self.module_sources[key]
            else:
                if self.library_root is None:
                    raise EvalError(f"cannot open module {key}: no library root")
                file = os.path.join(self.library_root, *path) + ".tuun"
                if not os.path.exists(file):
                    raise EvalError(f"cannot open module {key}: {file} not found")
                text = open(file).read()
            _, own = self.run_bindings(parse_module(text))
            self._modules[key] = own
        return self._modules[key]

    def run_bindings(self, bindings, env: Optional[Env] = None):
        """build_context (eval.rs:463-497): returns (environment, the module's OWN definitions —
        what an `open` of it exports; names it opened itself are not re-exported)."""
        env = env or self.prelude
        own: Dict[str, object] = {}
        for b in bindings:
            if b[0] == "open":
                env = env.extend(dict(self._module_exports(b[1])))
            else:
                frame: Dict[str, object] = {}
                _bind(b[1], evaluate(b[2], env.extend(_slider_bindings(b[3])) if b[3] else env), frame)
                env = env.extend(frame)
                own.update(frame)
        return env, own

    def evaluate_source(self, program: str, context: str = ""):
        """Evaluate one program expression after the bindings of `context` (module text)."""
        env, _ = self.run_bindings(parse_module(context))
        return evaluate(parse_program(program), env)

    def waveform(self, program: str, context: str = "", optimize: bool = True) -> Waveform:
        """parse -> evaluate -> (optimize): what a player hands to the generator."""
        from .optimizer import optimize as opt
        w = B.to_waveform(self.evaluate_source(program, context))
        return opt(w) if optimize else w
