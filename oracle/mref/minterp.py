"""Tree-walking interpreter for the MATLAB subset of mparse.py: runs the reference's UNMODIFIED
matlab_code/*.m files from where they lie (read-only) so that their outputs can pin the oracle.

TEST INFRASTRUCTURE (oracle/): never imported by the product package.

Value model (MATLAB semantics, 1-based, column-major, value copies):
  numeric / logical -> 2-D numpy array (float64 / bool_), scalars are 1x1
  char              -> Python str
  struct (array)    -> StructArr (1xN list of field dicts, shared ordered field list)
  cell              -> Cell (1xN list)
`sparse` matrices are held dense: every product the reference forms is mathematically identical,
only the summation order inside a product can differ (rounding level, far below the 1e-9 bar).
`rand` draws from an explicit stream set by the harness (Interp.rand_stream), so the reference's
own select_random_match.m runs unmodified and consumes exactly one uniform per executed hypothesis.
"""
import math
import os
import time

import numpy as np

from . import mparse


class MError(Exception):
    pass


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


class _Return(Exception):
    pass


class StructArr:
    """1xN struct array; N == 1 is a plain struct."""
    __slots__ = ("fields", "elems")

    def __init__(self, fields=None, elems=None):
        self.fields = list(fields or [])
        self.elems = elems if elems is not None else []

    def copy(self):
        return StructArr(self.fields, [dict(e) for e in self.elems])

    def __len__(self):
        return len(self.elems)

    def get(self, name):
        if len(self.elems) != 1:
            raise MError("field access %r on a %d-element struct array" % (name, len(self.elems)))
        if name not in self.elems[0]:
            raise MError("reference to non-existent field %r" % name)
        return self.elems[0][name]

    # convenience for Python harness code
    def __getitem__(self, i):
        return self.elems[i]


class Cell:
    __slots__ = ("items",)

    def __init__(self, items):
        self.items = list(items)


EMPTY = np.zeros((0, 0))


def mat(v):
    """Python/numpy value -> interpreter numeric value (2-D)."""
    if isinstance(v, np.ndarray):
        if v.ndim == 2:
            return v if v.dtype in (np.float64, np.bool_) else v.astype(np.float64)
        if v.ndim == 0:
            return v.reshape(1, 1).astype(np.float64)
        if v.ndim == 1:
            return v.reshape(-1, 1).astype(np.float64)
        raise MError("arrays with more than 2 dimensions are not supported")
    if isinstance(v, (bool, np.bool_)):
        return np.array([[bool(v)]])
    if isinstance(v, (int, float, np.integer, np.floating)):
        return np.array([[float(v)]])
    return v


def is_num(v):
    return isinstance(v, np.ndarray)


def fl(v):
    """numeric value as float64 (logical -> double; char -> codes)."""
    if isinstance(v, str):
        return np.array([[float(ord(c)) for c in v]]) if v else EMPTY
    if v.dtype == np.bool_:
        return v.astype(np.float64)
    return v


def scalar(v):
    v = fl(v)
    if v.size != 1:
        raise MError("scalar expected, got %s" % (v.shape,))
    return float(v.flat[0])


def truth(v):
    """MATLAB `if` semantics: non-empty and all elements non-zero."""
    if isinstance(v, str):
        return len(v) > 0
    if isinstance(v, (StructArr, Cell)):
        raise MError("struct/cell used as a condition")
    return v.size > 0 and bool(np.all(v != 0))


def is_empty(v):
    if isinstance(v, str):
        return len(v) == 0
    if isinstance(v, StructArr):
        return len(v.elems) == 0
    if isinstance(v, Cell):
        return len(v.items) == 0
    return v.size == 0


def size_of(v):
    if isinstance(v, str):
        return (1 if v else 0, len(v))
    if isinstance(v, StructArr):
        return (1 if v.elems else 0, len(v.elems))
    if isinstance(v, Cell):
        return (1 if v.items else 0, len(v.items))
    return v.shape


# ------------------------------------------------------------------------------------------
# indexing helpers
# ------------------------------------------------------------------------------------------
class _Colon:
    pass


COLON = _Colon()


def _index_vector(ix, dim_len):
    """index argument -> 0-based integer numpy vector (and its original shape)."""
    if ix is COLON:
        return np.arange(dim_len), None
    if ix.dtype == np.bool_:
        flat = ix.reshape(-1, order="F")
        if flat.size > dim_len and np.any(flat[dim_len:]):
            raise MError("logical index out of bounds")
        return np.nonzero(flat)[0], ix.shape
    flat = ix.reshape(-1, order="F")
    idx = flat.astype(np.int64)
    if np.any(idx != flat):
        raise MError("non-integer index %r" % (flat,))
    if np.any(idx < 1):
        raise MError("index must be a positive integer")
    return idx - 1, ix.shape


def index_numeric(a, args):
    if len(args) == 1:
        ix = args[0]
        if ix is COLON:
            return a.reshape(-1, 1, order="F")
        idx, ishape = _index_vector(ix, a.size)
        if idx.size and idx.max() >= a.size:
            raise MError("index %d out of bounds (numel %d)" % (idx.max() + 1, a.size))
        vals = a.reshape(-1, order="F")[idx]
        is_mask = ix.dtype == np.bool_
        a_vec = a.shape[0] == 1 or a.shape[1] == 1
        i_vec = ishape[0] == 1 or ishape[1] == 1
        if a_vec and (i_vec or is_mask) and a.size != 1:
            return vals.reshape(1, -1) if a.shape[0] == 1 else vals.reshape(-1, 1)
        if is_mask:
            return vals.reshape(-1, 1) if not (ishape[0] == 1) else vals.reshape(1, -1)
        if idx.size == 0 and a_vec:
            return np.zeros((1, 0)) if a.shape[0] == 1 and a.size != 1 else np.zeros(ishape)
        return vals.reshape(ishape, order="F")
    if len(args) == 2:
        r, _ = _index_vector(args[0], a.shape[0])
        c, _ = _index_vector(args[1], a.shape[1])
        if (r.size and r.max() >= a.shape[0]) or (c.size and c.max() >= a.shape[1]):
            raise MError("index out of bounds: (%s,%s) into %s" % (r.max() + 1 if r.size else 0,
                                                                  c.max() + 1 if c.size else 0, a.shape))
        return a[np.ix_(r, c)]
    raise MError("only 1- and 2-index subscripts are supported")


def assign_numeric(a, args, rhs):
    """a(args) = rhs with MATLAB growth rules; returns a new array."""
    rhs = fl(mat(rhs)) if not (is_num(rhs) and rhs.dtype == np.bool_ and a is not None and
                               a.dtype == np.bool_) else rhs
    if a is None:
        a = EMPTY
    if a.dtype == np.bool_ and rhs.dtype != np.bool_:
        a = a.astype(np.float64)
    if len(args) == 1:
        ix = args[0]
        if ix is COLON:
            return _assign_flat(a, np.arange(a.size), rhs)
        idx, _ = _index_vector(ix, a.size)
        if rhs.size == 0 and idx.size:          # deletion  a(idx) = []
            keep = np.ones(a.size, dtype=bool)
            keep[idx] = False
            vals = a.reshape(-1, order="F")[keep]
            return vals.reshape(1, -1) if a.shape[0] == 1 else vals.reshape(-1, 1)
        need = (idx.max() + 1) if idx.size else 0
        if need > a.size:
            if a.size == 0:
                a = np.zeros((1, need), dtype=a.dtype if a.size else np.float64)
            elif a.shape[0] == 1:
                a = np.hstack([a, np.zeros((1, need - a.size), dtype=a.dtype)])
            elif a.shape[1] == 1:
                a = np.vstack([a, np.zeros((need - a.size, 1), dtype=a.dtype)])
            else:
                raise MError("linear-index growth of a matrix")
        return _assign_flat(a, idx, rhs)
    if len(args) == 2:
        r_all, c_all = args[0] is COLON, args[1] is COLON
        if a.size == 0 and (r_all or c_all):
            # x(1,:) = v on an undefined / empty x: the colon dimension takes the size of rhs
            nr = rhs.shape[0] if r_all else None
            nc = rhs.shape[1] if c_all else None
            if r_all and not c_all:
                c, _ = _index_vector(args[1], 0)
                nr = rhs.shape[0] if rhs.shape[1] == c.size else rhs.size // max(c.size, 1)
                a = np.zeros((nr, 0))
            elif c_all and not r_all:
                r, _ = _index_vector(args[0], 0)
                nc = rhs.shape[1] if rhs.shape[0] == r.size else rhs.size // max(r.size, 1)
                a = np.zeros((0, nc))
        r, _ = _index_vector(args[0], a.shape[0])
        c, _ = _index_vector(args[1], a.shape[1])
        nr = max(a.shape[0], (r.max() + 1) if r.size else 0)
        nc = max(a.shape[1], (c.max() + 1) if c.size else 0)
        if (nr, nc) != a.shape:
            grown = np.zeros((nr, nc), dtype=a.dtype)
            grown[:a.shape[0], :a.shape[1]] = a
            a = grown
        else:
            a = a.copy()
        if rhs.size == 1:
            a[np.ix_(r, c)] = rhs.flat[0]
        else:
            if rhs.shape != (r.size, c.size):
                if rhs.size == r.size * c.size and (rhs.shape[0] == 1 or rhs.shape[1] == 1) and \
                        (r.size == 1 or c.size == 1):
                    rhs = rhs.reshape(r.size, c.size)
                else:
                    raise MError("subscripted assignment dimension mismatch: %s into (%d,%d)"
                                 % (rhs.shape, r.size, c.size))
            a[np.ix_(r, c)] = rhs
        return a
    raise MError("only 1- and 2-index assignments are supported")


def _assign_flat(a, idx, rhs):
    flat = a.reshape(-1, order="F").copy()
    if rhs.size == 1:
        flat[idx] = rhs.flat[0]
    else:
        if rhs.size != idx.size:
            raise MError("A(I) = B: number of elements differ (%d vs %d)" % (idx.size, rhs.size))
        flat[idx] = rhs.reshape(-1, order="F")
    return flat.reshape(a.shape, order="F")


def hcat(vals):
    vals = [v for v in vals if not (is_num(v) and v.size == 0)]
    if not vals:
        return EMPTY
    if all(isinstance(v, str) for v in vals):
        return "".join(vals)
    if any(isinstance(v, StructArr) for v in vals):
        out = None
        for v in vals:
            out = v.copy() if out is None else StructArr(out.fields, out.elems + [dict(e) for e in v.elems])
        return out
    vals = [fl(v) if not is_num(v) or v.dtype != np.bool_ else v for v in vals]
    if not all(v.dtype == np.bool_ for v in vals):
        vals = [fl(v) for v in vals]
    rows = vals[0].shape[0]
    for v in vals:
        if v.shape[0] != rows:
            raise MError("horizontal dimensions mismatch (%s)" % ", ".join(str(v.shape) for v in vals))
    return np.hstack(vals)


def vcat(vals):
    vals = [v for v in vals if not (is_num(v) and v.size == 0)]
    if not vals:
        return EMPTY
    if len(vals) == 1:
        return vals[0]
    if any(isinstance(v, str) for v in vals):
        raise MError("vertical concatenation of strings is not supported")
    if not all(v.dtype == np.bool_ for v in vals):
        vals = [fl(v) for v in vals]
    cols = vals[0].shape[1]
    for v in vals:
        if v.shape[1] != cols:
            raise MError("vertical dimensions mismatch (%s)" % ", ".join(str(v.shape) for v in vals))
    return np.vstack(vals)


# ------------------------------------------------------------------------------------------
# the interpreter
# ------------------------------------------------------------------------------------------
class Interp:
    def __init__(self, path, verbose=False):
        """path: list of directories searched in order for <name>.m (first hit wins)."""
        self.path = list(path)
        self.files = {}          # function name -> (main function dict, {local name: function dict})
        self.globals = {}
        self.rand_stream = None  # iterator of uniforms for rand()
        self.rand_drawn = 0
        self.verbose = verbose
        self.call_counts = {}    # name -> number of calls of path functions (evidence of what ran)
        self.sources_used = {}   # name -> file path
        self._tic = time.time()
        self.builtins = _make_builtins(self)

    # -- function lookup --------------------------------------------------------------------
    def find_function(self, name):
        if name in self.files:
            return self.files[name]
        for d in self.path:
            fn = os.path.join(d, name + ".m")
            if os.path.isfile(fn):
                with open(fn, "r", encoding="latin-1") as fh:
                    src = fh.read()
                funcs, script = mparse.parse_source(src.replace("\r\n", "\n").replace("\r", "\n"), fn)
                if not funcs:
                    entry = (dict(name=name, ins=[], outs=[], body=script, file=fn, script=True), {})
                else:
                    entry = (funcs[0], {f["name"]: f for f in funcs[1:]})
                self.files[name] = entry
                self.sources_used[name] = fn
                return entry
        self.files[name] = None
        return None

    def call(self, name, *args, nargout=1):
        """Python entry: call a .m function / builtin by name with Python/numpy values."""
        outs = self.call_function(name, [mat(a) for a in args], nargout, None)
        if nargout == 1:
            return outs[0]
        return outs

    def call_function(self, name, args, nargout, local_funcs):
        if local_funcs and name in local_funcs:
            return self.run_function(local_funcs[name], args, nargout, local_funcs)
        entry = self.find_function(name)
        if entry is not None:
            main, locs = entry
            self.call_counts[name] = self.call_counts.get(name, 0) + 1
            return self.run_function(main, args, nargout, locs)
        b = self.builtins.get(name)
        if b is not None:
            out = b(args, nargout)
            return out if isinstance(out, list) else [out]
        raise MError("undefined function or variable %r" % name)

    def run_function(self, f, args, nargout, local_funcs):
        ws = {}
        ins = f["ins"]
        if ins and ins[-1] == "varargin":
            fixed = ins[:-1]
            ws["varargin"] = Cell(args[len(fixed):])
            for n_, a in zip(fixed, args):
                ws[n_] = a
        else:
            if len(args) > len(ins):
                raise MError("%s: too many input arguments (%d > %d)" % (f["name"], len(args), len(ins)))
            for n_, a in zip(ins, args):
                ws[n_] = a
        ws["nargin"] = mat(len(args))
        ws["nargout"] = mat(nargout)
        frame = _Frame(ws, local_funcs, f)
        try:
            self.exec_block(f["body"], frame)
        except _Return:
            pass
        outs = []
        for k, o in enumerate(f["outs"][:max(nargout, 1)]):
            if o in frame.global_names:
                outs.append(self.globals[o])
            elif o in ws:
                outs.append(ws[o])
            else:
                if k < nargout:
                    raise MError("%s: output %r not assigned" % (f["name"], o))
        return outs

    # -- statements -------------------------------------------------------------------------
    def exec_block(self, body, fr):
        for st in body:
            self.exec_stmt(st, fr)

    def exec_stmt(self, st, fr):
        kind = st[0]
        if kind == "assign":
            _, lvs, rhs, _q = st
            if len(lvs) == 1:
                val = self.eval(rhs, fr)
                self.assign(lvs[0], val, fr)
            else:
                vals = self.eval_multi(rhs, fr, len(lvs))
                if len(vals) < len([lv for lv in lvs if lv is not None]):
                    raise MError("not enough output values")
                for lv, v in zip(lvs, vals):
                    if lv is not None:
                        self.assign(lv, v, fr)
        elif kind == "expr":
            e = st[1]
            if e[0] == "ref" and not e[2] and e[1] not in fr.ws and e[1] not in fr.global_names:
                self.call_ref(e[1], [], 0, fr)      # bare command / call without outputs
            elif e[0] == "ref" and e[1] not in fr.ws and e[1] not in fr.global_names:
                self.eval_multi(e, fr, 0)
            else:
                fr.ws["ans"] = self.eval(e, fr)
        elif kind == "if":
            for cond, body in st[1]:
                if truth(self.eval(cond, fr)):
                    self.exec_block(body, fr)
                    return
            if st[2] is not None:
                self.exec_block(st[2], fr)
        elif kind == "for":
            _, var, e, body = st
            rng = self.eval(e, fr)
            if isinstance(rng, str):
                rng = fl(rng)
            ncols = rng.shape[1] if rng.size else 0
            for j in range(ncols):
                col = rng[:, j:j + 1]
                self.set_var(var, col.copy(), fr)
                try:
                    self.exec_block(body, fr)
                except _Break:
                    break
                except _Continue:
                    continue
        elif kind == "while":
            while truth(self.eval(st[1], fr)):
                try:
                    self.exec_block(st[2], fr)
                except _Break:
                    break
                except _Continue:
                    continue
        elif kind == "break":
            raise _Break()
        elif kind == "continue":
            raise _Continue()
        elif kind == "return":
            raise _Return()
        elif kind == "global":
            for n_ in st[1]:
                fr.global_names.add(n_)
                self.globals.setdefault(n_, EMPTY)
        else:
            raise MError("unknown statement %r" % (kind,))

    def get_var(self, name, fr):
        if name in fr.global_names:
            return self.globals[name]
        return fr.ws.get(name)

    def set_var(self, name, val, fr):
        if name in fr.global_names:
            self.globals[name] = val
        else:
            fr.ws[name] = val

    def assign(self, lv, val, fr):
        _, name, acc = lv
        if not acc:
            self.set_var(name, val, fr)
            return
        base = self.get_var(name, fr)
        self.set_var(name, self.assign_into(base, acc, val, fr), fr)

    def assign_into(self, base, acc, rhs, fr):
        if not acc:
            return rhs
        kind, arg = acc[0]
        rest = acc[1:]
        if kind == ".":
            if base is None or (is_num(base) and base.size == 0):
                base = StructArr([], [{}])
            if not isinstance(base, StructArr):
                raise MError("field assignment to a non-struct value")
            if len(base.elems) != 1:
                raise MError("field assignment to a %d-element struct array" % len(base.elems))
            new = base.copy()
            old = new.elems[0].get(arg)
            new.elems[0][arg] = self.assign_into(old, rest, rhs, fr)
            if arg not in new.fields:
                new.fields.append(arg)
            return new
        if kind == "()":
            struct_target = isinstance(base, StructArr) or \
                ((base is None or (is_num(base) and base.size == 0)) and
                 ((rest and rest[0][0] == ".") or isinstance(rhs, StructArr)))
            if struct_target:
                if base is None or is_num(base):
                    base = StructArr([], [])
                args = self.eval_index_args(arg, base, fr)
                if len(args) == 2:      # features_info(1, i)
                    if scalar(args[0]) != 1:
                        raise MError("struct arrays are 1xN")
                    args = [args[1]]
                i = int(scalar(args[0]))
                new = base.copy()
                while len(new.elems) < i:
                    new.elems.append({f: EMPTY for f in new.fields})
                if rest:
                    elem = StructArr(new.fields, [new.elems[i - 1]])
                    res = self.assign_into(elem, rest, rhs, fr)
                else:
                    if not isinstance(rhs, StructArr) or len(rhs.elems) != 1:
                        raise MError("struct element assignment needs a scalar struct")
                    res = rhs
                for f in res.fields:
                    if f not in new.fields:
                        new.fields.append(f)
                        for e in new.elems:
                            e.setdefault(f, EMPTY)
                e = dict(res.elems[0])
                for f in new.fields:
                    e.setdefault(f, EMPTY)
                new.elems[i - 1] = e
                return new
            if rest:
                raise MError("chained assignment after () on a numeric value")
            if isinstance(base, str):
                base = fl(base)
            args = self.eval_index_args(arg, base if base is not None else EMPTY, fr)
            return assign_numeric(base, args, rhs)
        if kind == "{}":
            if base is None or (is_num(base) and base.size == 0):
                base = Cell([])
            args = self.eval_index_args(arg, EMPTY, fr)
            i = int(scalar(args[-1]))
            items = list(base.items)
            while len(items) < i:
                items.append(EMPTY)
            items[i - 1] = self.assign_into(items[i - 1], rest, rhs, fr)
            return Cell(items)
        raise MError("bad l-value")

    # -- expressions ------------------------------------------------------------------------
    def eval_multi(self, e, fr, nargout):
        """Evaluates a call expression asking for nargout outputs; returns a list."""
        if e[0] == "ref" and e[1] not in fr.ws and e[1] not in fr.global_names:
            name, acc = e[1], e[2]
            if acc and acc[0][0] == "()":
                args = [self.eval(a, fr) for a in acc[0][1]]
                outs = self.call_ref(name, args, nargout, fr)
                if len(acc) > 1:
                    outs = [self.apply_accessors(outs[0], acc[1:], fr)]
                return outs
            if not acc:
                return self.call_ref(name, [], nargout, fr)
        return [self.eval(e, fr)]

    def call_ref(self, name, args, nargout, fr):
        return self.call_function(name, args, nargout, fr.local_funcs)

    def eval(self, e, fr):
        k = e[0]
        if k == "num":
            return np.array([[e[1]]])
        if k == "str":
            return e[1]
        if k == "paren":
            return self.eval(e[1], fr)
        if k == "ref":
            name, acc = e[1], e[2]
            val = self.get_var(name, fr)
            if val is not None:
                return self.apply_accessors(val, acc, fr) if acc else val
            outs = self.eval_multi(e, fr, 1)
            if not outs:
                raise MError("%s returned nothing" % name)
            return outs[0]
        if k == "bin":
            op = e[1]
            if op == "&&":
                a = self.eval(e[2], fr)
                if not truth(a):
                    return np.array([[False]])
                return np.array([[truth(self.eval(e[3], fr))]])
            if op == "||":
                a = self.eval(e[2], fr)
                if truth(a):
                    return np.array([[True]])
                return np.array([[truth(self.eval(e[3], fr))]])
            return binop(op, self.eval(e[2], fr), self.eval(e[3], fr))
        if k == "un":
            v = self.eval(e[2], fr)
            if e[1] == "-":
                return -fl(v)
            if e[1] == "+":
                return fl(v)
            return fl(v) == 0
        if k == "post":
            v = self.eval(e[2], fr)
            if isinstance(v, str):
                raise MError("transpose of a string")
            return v.T
        if k == "range":
            a = scalar(self.eval(e[1], fr))
            b = scalar(self.eval(e[3], fr))
            s = scalar(self.eval(e[2], fr)) if e[2] is not None else 1.0
            return make_range(a, s, b)
        if k == "matrix":
            rows = []
            for row in e[1]:
                rows.append(hcat([self.eval(x, fr) for x in row]))
            if len(rows) == 1:
                return rows[0]
            return vcat(rows)
        if k == "cell":
            items = []
            for row in e[1]:
                items.extend(self.eval(x, fr) for x in row)
            return Cell(items)
        if k == "end":
            if not fr.end_stack:
                raise MError("`end` outside an index expression")
            val, pos, n = fr.end_stack[-1]
            sz = size_of(val)
            if n == 1:
                return mat(int(np.prod(sz)))
            return mat(sz[pos] if pos < 2 else 1)
        if k == "colon":
            return COLON
        raise MError("unknown expression %r" % (k,))

    def eval_index_args(self, arg_exprs, base, fr):
        out = []
        n = len(arg_exprs)
        for pos, a in enumerate(arg_exprs):
            if a[0] == "colon":
                out.append(COLON)
                continue
            fr.end_stack.append((base, pos, n))
            try:
                v = self.eval(a, fr)
            finally:
                fr.end_stack.pop()
            if isinstance(v, str) and v == ":":
                v = COLON
            out.append(v)
        return out

    def apply_accessors(self, val, acc, fr):
        for kind, arg in acc:
            if kind == ".":
                if not isinstance(val, StructArr):
                    raise MError("field access %r on a non-struct" % arg)
                val = val.get(arg)
            elif kind == "()":
                args = self.eval_index_args(arg, val, fr)
                if isinstance(val, StructArr):
                    if len(args) == 2:
                        args = [args[1]]
                    if args[0] is COLON:
                        continue
                    idx, _ = _index_vector(args[0], len(val.elems))
                    if idx.size and idx.max() >= len(val.elems):
                        raise MError("index exceeds struct array bounds (%d > %d)"
                                     % (idx.max() + 1, len(val.elems)))
                    val = StructArr(val.fields, [val.elems[i] for i in idx])
                elif isinstance(val, str):
                    r = index_numeric(fl(val), args)
                    val = "".join(chr(int(c)) for c in r.reshape(-1))
                elif isinstance(val, Cell):
                    idx, _ = _index_vector(args[-1], len(val.items))
                    val = Cell([val.items[i] for i in idx])
                else:
                    val = index_numeric(val, args)
            elif kind == "{}":
                if not isinstance(val, Cell):
                    raise MError("{} on a non-cell")
                args = self.eval_index_args(arg, mat(np.zeros((1, len(val.items)))), fr)
                i = int(scalar(args[-1]))
                if i > len(val.items):
                    raise MError("index exceeds cell bounds")
                val = val.items[i - 1]
        return val


class _Frame:
    __slots__ = ("ws", "local_funcs", "func", "global_names", "end_stack")

    def __init__(self, ws, local_funcs, func):
        self.ws = ws
        self.local_funcs = local_funcs
        self.func = func
        self.global_names = set()
        self.end_stack = []


def make_range(a, s, b):
    if s == 0 or (s > 0 and a > b) or (s < 0 and a < b):
        return np.zeros((1, 0))
    n = int(math.floor((b - a) / s * (1 + 1e-15) + 1e-10)) + 1
    return (a + s * np.arange(n, dtype=np.float64)).reshape(1, -1)


def _bcast(a, b, what):
    if a.shape == b.shape or a.size == 1 or b.size == 1:
        return
    # implicit expansion (R2016b+ / Octave): allowed when singleton dims match up
    for da, db in zip(a.shape, b.shape):
        if da != db and da != 1 and db != 1:
            raise MError("%s: nonconformant arguments (%s vs %s)" % (what, a.shape, b.shape))


def binop(op, a, b):
    if op in ("==", "~=") and isinstance(a, str) and isinstance(b, str) and len(a) != len(b):
        raise MError("comparison of strings of different length")
    a, b = fl(a), fl(b)
    if op == "*":
        if a.size == 1 or b.size == 1:
            return a * b
        if a.shape[1] != b.shape[0]:
            raise MError("*: nonconformant arguments (%s * %s)" % (a.shape, b.shape))
        return a @ b
    if op == "/":
        if b.size == 1:
            return a / b
        return np.linalg.solve(b.T, a.T).T
    if op == "\\":
        if a.size == 1:
            return b / a
        return np.linalg.solve(a, b)
    if op == "^":
        if a.size == 1 and b.size == 1:
            return _pow(a, b)
        if b.size == 1 and a.shape[0] == a.shape[1] and float(b.flat[0]) == int(b.flat[0]):
            return np.linalg.matrix_power(a, int(b.flat[0]))
        raise MError("matrix power is not supported")
    _bcast(a, b, op)
    if op == "+":
        return a + b
    if op == "-":
        return a - b
    if op == ".*":
        return a * b
    if op == "./":
        return a / b
    if op == ".\\":
        return b / a
    if op == ".^":
        return _pow(a, b)
    if op == "==":
        return a == b
    if op == "~=":
        return a != b
    if op == "<":
        return a < b
    if op == "<=":
        return a <= b
    if op == ">":
        return a > b
    if op == ">=":
        return a >= b
    if op == "&":
        return (a != 0) & (b != 0)
    if op == "|":
        return (a != 0) | (b != 0)
    raise MError("unknown operator %r" % op)


def _pow(a, b):
    # x.^2 is computed as x*x by MATLAB/Octave (exactly rounded product); general powers go to pow()
    if b.size == 1:
        p = float(b.flat[0])
        if p == 2.0:
            return a * a
        if p == 3.0:
            return np.power(a, 3.0)
    return np.power(a, b)


# ------------------------------------------------------------------------------------------
# builtins
# ------------------------------------------------------------------------------------------
def _dims(args):
    if len(args) == 0:
        return (1, 1)
    if len(args) == 1:
        a = fl(args[0])
        if a.size == 1:
            n = int(a.flat[0])
            return (n, n)
        return tuple(int(x) for x in a.reshape(-1))
    return tuple(int(scalar(a)) for a in args)


def _elementwise(fn):
    def f(args, nargout):
        return fn(fl(args[0]))
    return f


def _reduce(fn, logical=False):
    def f(args, nargout):
        a = fl(args[0])
        if len(args) > 1:
            dim = int(scalar(args[1]))
            r = fn(a, axis=dim - 1, keepdims=True) if a.size else np.zeros((1, 1))
            return r
        if a.size == 0:
            return fn(np.zeros(0)).reshape(1, 1)
        if a.shape[0] == 1 or a.shape[1] == 1:
            return np.asarray(fn(a.reshape(-1))).reshape(1, 1)
        return np.asarray(fn(a, axis=0)).reshape(1, -1)
    return f


def _make_builtins(interp):
    B = {}

    def const(v):
        return lambda args, nargout: mat(v)

    B["pi"] = const(math.pi)
    B["Inf"] = B["inf"] = const(float("inf"))
    B["NaN"] = B["nan"] = const(float("nan"))
    B["true"] = lambda a, n: np.array([[True]])
    B["false"] = lambda a, n: np.array([[False]])

    def b_eps(args, nargout):
        if args:
            return np.spacing(np.abs(fl(args[0])))
        return mat(np.finfo(np.float64).eps)
    B["eps"] = b_eps

    B["zeros"] = lambda a, n: np.zeros(_dims(a))
    B["ones"] = lambda a, n: np.ones(_dims(a))

    def b_eye(args, nargout):
        d = _dims(args)
        return np.eye(d[0], d[1])
    B["eye"] = B["speye"] = b_eye

    def b_sparse(args, nargout):
        if len(args) == 1:
            return fl(args[0])
        if len(args) == 2:
            return np.zeros(_dims(args))
        raise MError("sparse(i,j,v) is not supported")
    B["sparse"] = b_sparse
    B["full"] = lambda a, n: fl(a[0])
    B["double"] = lambda a, n: fl(a[0]) if not isinstance(a[0], (StructArr, Cell)) else a[0]
    B["single"] = B["double"]
    B["logical"] = lambda a, n: fl(a[0]) != 0
    B["inv"] = lambda a, n: np.linalg.inv(fl(a[0])) if fl(a[0]).size > 1 else 1.0 / fl(a[0])
    B["det"] = lambda a, n: mat(np.linalg.det(fl(a[0])))
    B["trace"] = lambda a, n: mat(np.trace(fl(a[0])))
    B["transpose"] = lambda a, n: fl(a[0]).T

    def b_eig(args, nargout):
        a = fl(args[0])
        if nargout > 1:
            raise MError("[V,D]=eig is not supported")
        if np.array_equal(a, a.T):
            return np.linalg.eigvalsh(a).reshape(-1, 1)
        w = np.linalg.eigvals(a)
        if np.all(np.isreal(w)):
            w = np.sort(w.real)
        return w.reshape(-1, 1)
    B["eig"] = b_eig

    def b_chol(args, nargout):
        return np.linalg.cholesky(fl(args[0])).T
    B["chol"] = b_chol

    def b_diag(args, nargout):
        a = fl(args[0])
        if a.shape[0] == 1 or a.shape[1] == 1:
            return np.diag(a.reshape(-1))
        return np.diag(a).reshape(-1, 1)
    B["diag"] = b_diag

    def b_norm(args, nargout):
        a = fl(args[0])
        if a.size == 0:
            return mat(0.0)
        if a.shape[0] == 1 or a.shape[1] == 1:
            v = a.reshape(-1)
            if len(args) > 1:
                p = args[1]
                if isinstance(p, str):
                    return mat(np.linalg.norm(v, np.inf if p.lower() == "inf" else None))
                return mat(np.linalg.norm(v, scalar(p)))
            return mat(np.linalg.norm(v))
        if len(args) > 1 and isinstance(args[1], str) and args[1] == "fro":
            return mat(np.linalg.norm(a, "fro"))
        return mat(np.linalg.norm(a, 2))
    B["norm"] = b_norm

    for name, fn in (("sqrt", np.sqrt), ("sin", np.sin), ("cos", np.cos), ("tan", np.tan),
                     ("asin", np.arcsin), ("acos", np.arccos), ("atan", np.arctan), ("exp", np.exp),
                     ("log", np.log), ("abs", np.abs), ("floor", np.floor), ("ceil", np.ceil),
                     ("fix", np.trunc), ("sign", np.sign)):
        B[name] = _elementwise(fn)
    # MATLAB round: half away from zero
    B["round"] = _elementwise(lambda a: np.sign(a) * np.floor(np.abs(a) + 0.5))
    B["atan2"] = lambda a, n: np.arctan2(fl(a[0]), fl(a[1]))
    B["mod"] = lambda a, n: np.mod(fl(a[0]), fl(a[1]))
    B["rem"] = lambda a, n: np.fmod(fl(a[0]), fl(a[1]))
    B["sum"] = _reduce(np.sum)
    B["prod"] = _reduce(np.prod)
    B["mean"] = _reduce(np.mean)
    B["all"] = lambda a, n: np.array([[bool(np.all(fl(a[0]) != 0))]]) \
        if min(fl(a[0]).shape + (1,)) <= 1 or fl(a[0]).size == 0 else np.all(fl(a[0]) != 0, axis=0, keepdims=True)
    B["any"] = lambda a, n: np.array([[bool(np.any(fl(a[0]) != 0))]]) \
        if min(fl(a[0]).shape + (1,)) <= 1 or fl(a[0]).size == 0 else np.any(fl(a[0]) != 0, axis=0, keepdims=True)
    B["not"] = lambda a, n: fl(a[0]) == 0

    def minmax(fn, fn2, argfn):
        def f(args, nargout):
            if len(args) >= 2 and not (is_num(args[1]) and args[1].size == 0):
                return fn2(fl(args[0]), fl(args[1]))
            a = fl(args[0])
            if a.size == 0:
                return [EMPTY, EMPTY][:max(nargout, 1)]
            if a.shape[0] == 1 or a.shape[1] == 1:
                v = a.reshape(-1)
                return [mat(fn(v)), mat(int(argfn(v)) + 1)][:max(nargout, 1)]
            return [fn(a, axis=0).reshape(1, -1), (argfn(a, axis=0) + 1.0).reshape(1, -1)][:max(nargout, 1)]
        return f
    B["max"] = minmax(np.max, np.maximum, np.argmax)
    B["min"] = minmax(np.min, np.minimum, np.argmin)

    def b_find(args, nargout):
        a = fl(args[0])
        idx = np.nonzero(a.reshape(-1, order="F"))[0] + 1.0
        if a.shape[0] == 1 and a.size != 1:
            return idx.reshape(1, -1)
        return idx.reshape(-1, 1)
    B["find"] = b_find

    def b_length(args, nargout):
        sz = size_of(args[0])
        return mat(0 if min(sz) == 0 else max(sz))
    B["length"] = b_length
    B["numel"] = lambda a, n: mat(int(np.prod(size_of(a[0]))))

    def b_size(args, nargout):
        sz = size_of(args[0])
        if len(args) > 1:
            d = int(scalar(args[1]))
            return mat(sz[d - 1] if d <= 2 else 1)
        if nargout <= 1:
            return np.array([[float(sz[0]), float(sz[1])]])
        return [mat(sz[0]), mat(sz[1])] + [mat(1)] * (nargout - 2)
    B["size"] = b_size
    B["isempty"] = lambda a, n: np.array([[is_empty(a[0])]])
    B["isfield"] = lambda a, n: np.array([[isinstance(a[0], StructArr) and a[1] in a[0].fields]])
    B["isstruct"] = lambda a, n: np.array([[isinstance(a[0], StructArr)]])
    B["isa"] = lambda a, n: np.array([[(a[1] == "double" and is_num(a[0]) and a[0].dtype == np.float64) or
                                       (a[1] == "struct" and isinstance(a[0], StructArr)) or
                                       (a[1] == "char" and isinstance(a[0], str))]])
    B["ischar"] = lambda a, n: np.array([[isinstance(a[0], str)]])

    def b_reshape(args, nargout):
        a = fl(args[0])
        d = list(_dims(args[1:]))
        return a.reshape(-1, order="F").reshape(d, order="F")
    B["reshape"] = b_reshape

    def b_repmat(args, nargout):
        d = _dims(args[1:])
        return np.tile(fl(args[0]), d)
    B["repmat"] = b_repmat

    def b_cross(args, nargout):
        a, b = fl(args[0]), fl(args[1])
        r = np.cross(a.reshape(-1), b.reshape(-1))
        return r.reshape(a.shape)
    B["cross"] = b_cross
    B["dot"] = lambda a, n: mat(float(np.dot(fl(a[0]).reshape(-1), fl(a[1]).reshape(-1))))
    B["kron"] = lambda a, n: np.kron(fl(a[0]), fl(a[1]))

    B["strcmp"] = lambda a, n: np.array([[isinstance(a[0], str) and isinstance(a[1], str) and a[0] == a[1]]])

    def b_strncmp(args, nargout):
        s1, s2, k = args[0], args[1], int(scalar(args[2]))
        ok = isinstance(s1, str) and isinstance(s2, str) and len(s1) >= k and len(s2) >= k and s1[:k] == s2[:k]
        return np.array([[ok]])
    B["strncmp"] = b_strncmp

    def b_rand(args, nargout):
        d = _dims(args)
        if interp.rand_stream is None:
            raise MError("rand() called but no uniform stream was supplied by the harness")
        out = np.zeros(d)
        flat = out.reshape(-1, order="F")      # column-major fill order, like MATLAB
        vals = []
        for _ in range(flat.size):
            try:
                vals.append(next(interp.rand_stream))
            except StopIteration:
                raise MError("uniform stream exhausted after %d draws" % interp.rand_drawn)
            interp.rand_drawn += 1
        return np.array(vals).reshape(d, order="F")
    B["rand"] = b_rand

    def b_struct(args, nargout):
        s = StructArr([], [{}])
        for k in range(0, len(args), 2):
            s.fields.append(args[k])
            s.elems[0][args[k]] = args[k + 1]
        return s
    B["struct"] = b_struct

    def b_error(args, nargout):
        raise MError("error(): " + " ".join(str(a) for a in args))
    B["error"] = b_error

    def quiet(args, nargout):
        return []
    for name in ("fprintf", "printf", "disp", "display", "warning", "clc", "close", "figure", "drawnow",
                 "hold", "addpath", "rng", "format"):
        B[name] = quiet

    def b_tic(args, nargout):
        interp._tic = time.time()
        return [mat(interp._tic)] if nargout else []
    B["tic"] = b_tic
    B["toc"] = lambda a, n: mat(time.time() - (scalar(a[0]) if a else interp._tic))
    B["squeeze"] = lambda a, n: a[0]
    B["isreal"] = lambda a, n: np.array([[True]])
    B["issparse"] = lambda a, n: np.array([[False]])
    B["nnz"] = lambda a, n: mat(int(np.count_nonzero(fl(a[0]))))
    B["cumsum"] = lambda a, n: np.cumsum(fl(a[0]), axis=0 if fl(a[0]).shape[0] > 1 else 1)
    return B
