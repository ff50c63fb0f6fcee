"""Lexer + recursive-descent parser for the MATLAB subset the reference's matlab_code/*.m files use.

TEST INFRASTRUCTURE (part of oracle/): this exists so that the UNMODIFIED reference sources under
/root/reference/matlab_code can be *executed* here (no Octave/MATLAB in the image) and pin the
hand-written oracle.  Nothing under ekf-slam_b200/ imports it.

Covered syntax: functions (several per file, optional closing `end`, varargin), scripts,
if/elseif/else, for, while, break/continue/return, global, multi-assignment, struct-array /
field / cell l-values, matrix literals with MATLAB's whitespace rules ([a -b] vs [a - b], newline
and `;` row separators, `...` continuation), transpose vs string quotes, `end` inside indices,
ranges, short-circuit and element-wise operators, comments.  Not covered (unused by the reference
path): switch, try, classes, function handles, nested functions, command syntax.

AST = nested tuples:
  expressions  ('num', v) ('str', s) ('colon',) ('end',) ('ref', name, accessors)
               ('bin', op, a, b) ('un', op, a) ('post', op, a) ('range', a, step|None, b)
               ('matrix', rows) ('cell', rows)          accessors: ('()', args) ('.', name) ('{}', args)
  statements   ('expr', e, quiet) ('assign', lvalues, e, quiet) ('if', [(cond, body)], else_body)
               ('for', var, e, body) ('while', cond, body) ('break',) ('continue',) ('return',)
               ('global', names)
  function     dict(name=, ins=, outs=, body=, file=)
"""
import re

KEYWORDS = {"function", "end", "if", "elseif", "else", "for", "while", "break", "continue", "return",
            "global", "switch", "case", "otherwise", "try", "catch"}

_num_re = re.compile(r"(\d+(\.(?![*/^\\'])\d*)?|\.\d+)([eEdD][+-]?\d+)?")
_id_re = re.compile(r"[A-Za-z_][A-Za-z0-9_]*")
_ops3 = ("...",)
_ops2 = ("==", "~=", "<=", ">=", "&&", "||", ".*", "./", ".^", ".'", ".\\")
_ops1 = "+-*/\\^<>=&|~:(),;[]{}.'@!"


class Tok:
    __slots__ = ("kind", "val", "ws", "line")

    def __init__(self, kind, val, ws, line):
        self.kind, self.val, self.ws, self.line = kind, val, ws, line

    def __repr__(self):
        return "Tok(%s,%r,ws=%d,l%d)" % (self.kind, self.val, self.ws, self.line)


class MSyntaxError(Exception):
    pass


def tokenize(src, fname="<m>"):
    """Token stream.  Token.ws = 1 when whitespace (or a continuation) precedes the token.
    kinds: num str id kw op nl eof; `end` inside ()/{}/[] is emitted as ('op','end')."""
    toks = []
    i, n, line = 0, len(src), 1
    stack = []      # open delimiters
    ws = 0
    while i < n:
        c = src[i]
        if c in " \t\r":
            i += 1
            ws = 1
            continue
        if c == "%":
            while i < n and src[i] != "\n":
                i += 1
            continue
        if src.startswith("...", i):
            while i < n and src[i] != "\n":
                i += 1
            i += 1
            line += 1
            ws = 1
            continue
        if c == "\n":
            toks.append(Tok("nl", "\n", ws, line))
            i += 1
            line += 1
            ws = 0
            continue
        if c == "'":
            prev = toks[-1] if toks else None
            is_transpose = (prev is not None and
                            (prev.kind in ("num", "id") or
                             (prev.kind == "op" and prev.val in (")", "]", "}", "'", ".'", "end"))))
            if is_transpose and ws and stack and stack[-1] in "[{":
                is_transpose = False     # [a 'str'] : whitespace makes it a new element
            if is_transpose:
                toks.append(Tok("op", "'", ws, line))
                i += 1
                ws = 0
                continue
            j = i + 1
            buf = []
            while True:
                if j >= n or src[j] == "\n":
                    raise MSyntaxError("%s:%d: unterminated string" % (fname, line))
                if src[j] == "'":
                    if j + 1 < n and src[j + 1] == "'":
                        buf.append("'")
                        j += 2
                        continue
                    break
                buf.append(src[j])
                j += 1
            toks.append(Tok("str", "".join(buf), ws, line))
            i = j + 1
            ws = 0
            continue
        if c == '"':
            j = src.index('"', i + 1)
            toks.append(Tok("str", src[i + 1:j], ws, line))
            i = j + 1
            ws = 0
            continue
        m = _num_re.match(src, i)
        if m and (c.isdigit() or (c == "." and i + 1 < n and src[i + 1].isdigit())):
            txt = m.group(0).replace("d", "e").replace("D", "e")
            toks.append(Tok("num", float(txt), ws, line))
            i = m.end()
            ws = 0
            continue
        m = _id_re.match(src, i)
        if m:
            word = m.group(0)
            if word == "end" and stack:
                toks.append(Tok("op", "end", ws, line))
            elif word in KEYWORDS and not (toks and toks[-1].kind == "op" and toks[-1].val == "."):
                toks.append(Tok("kw", word, ws, line))
            else:
                toks.append(Tok("id", word, ws, line))
            i = m.end()
            ws = 0
            continue
        two = src[i:i + 2]
        if two in _ops2:
            toks.append(Tok("op", two, ws, line))
            i += 2
            ws = 0
            continue
        if c in _ops1:
            if c in "([{":
                stack.append(c)
            elif c in ")]}":
                if not stack:
                    raise MSyntaxError("%s:%d: unbalanced %s" % (fname, line, c))
                stack.pop()
            toks.append(Tok("op", c, ws, line))
            i += 1
            ws = 0
            continue
        raise MSyntaxError("%s:%d: unexpected character %r" % (fname, line, c))
    toks.append(Tok("nl", "\n", ws, line))
    toks.append(Tok("eof", None, 0, line))
    return toks


class Parser:
    def __init__(self, src, fname="<m>"):
        self.fname = fname
        self.t = tokenize(src, fname)
        self.p = 0

    # -- token helpers ----------------------------------------------------------------------
    def peek(self, k=0):
        return self.t[min(self.p + k, len(self.t) - 1)]

    def next(self):
        tok = self.t[self.p]
        self.p += 1
        return tok

    def is_op(self, val, k=0):
        tok = self.peek(k)
        return tok.kind == "op" and tok.val == val

    def is_kw(self, val, k=0):
        tok = self.peek(k)
        return tok.kind == "kw" and tok.val == val

    def expect_op(self, val):
        tok = self.next()
        if tok.kind != "op" or tok.val != val:
            self.err("expected %r, got %r" % (val, tok.val), tok)
        return tok

    def err(self, msg, tok=None):
        tok = tok or self.peek()
        raise MSyntaxError("%s:%d: %s" % (self.fname, tok.line, msg))

    def skip_seps(self):
        while self.peek().kind == "nl" or self.is_op(";") or self.is_op(","):
            self.next()

    # -- file level -------------------------------------------------------------------------
    def parse_file(self):
        """Returns (functions, script_body).  A function file has functions and an empty script."""
        self.skip_seps()
        funcs, script = [], []
        if self.is_kw("function"):
            while self.is_kw("function"):
                funcs.append(self.parse_function())
                self.skip_seps()
            if self.peek().kind != "eof":
                self.err("trailing code after functions")
        else:
            script = self.parse_block(("eof",))
        return funcs, script

    def parse_function(self):
        self.next()  # function
        outs = []
        # forms: function name(...) | function out = name(...) | function [o1,o2] = name(...)
        if self.is_op("["):
            self.next()
            while not self.is_op("]"):
                if self.is_op(","):
                    self.next()
                    continue
                outs.append(self.next().val)
            self.next()
            self.expect_op("=")
            name = self.next().val
        else:
            first = self.next().val
            if self.is_op("="):
                self.next()
                outs = [first]
                name = self.next().val
            else:
                name = first
        ins = []
        if self.is_op("("):
            self.next()
            while not self.is_op(")"):
                if self.is_op(","):
                    self.next()
                    continue
                tok = self.next()
                ins.append("~" if tok.val == "~" else tok.val)
            self.next()
        body = self.parse_block(("function", "end", "eof"))
        if self.is_kw("end"):
            self.next()
        return dict(name=name, ins=ins, outs=outs, body=body, file=self.fname)

    # -- statements -------------------------------------------------------------------------
    def parse_block(self, stops):
        body = []
        while True:
            self.skip_seps()
            tok = self.peek()
            if tok.kind == "eof":
                if "eof" in stops:
                    return body
                self.err("unexpected end of file")
            if tok.kind == "kw" and tok.val in stops:
                return body
            body.append(self.parse_statement())

    def end_of_statement(self):
        """Consumes the terminator; returns True when output is suppressed by `;`."""
        quiet = False
        tok = self.peek()
        if tok.kind == "op" and tok.val == ";":
            quiet = True
            self.next()
        elif tok.kind == "op" and tok.val == ",":
            self.next()
        elif tok.kind == "nl":
            self.next()
        elif tok.kind == "eof" or tok.kind == "kw":
            pass   # `if x break; end` : the next statement follows directly
        elif tok.kind == "id":
            pass   # `if cond stmt` on one line
        else:
            self.err("unexpected %r after statement" % (tok.val,))
        return quiet

    def parse_statement(self):
        tok = self.peek()
        if tok.kind == "kw":
            kw = tok.val
            if kw == "if":
                return self.parse_if()
            if kw == "for":
                return self.parse_for()
            if kw == "while":
                self.next()
                cond = self.parse_expr()
                body = self.parse_block(("end",))
                self.next()
                return ("while", cond, body)
            if kw in ("break", "continue", "return"):
                self.next()
                self.end_of_statement()
                return (kw,)
            if kw == "global":
                self.next()
                names = []
                while self.peek().kind == "id":
                    names.append(self.next().val)
                self.end_of_statement()
                return ("global", names)
            self.err("unsupported keyword %r" % kw)
        # assignment?
        save = self.p
        lvs = self.try_parse_lhs()
        if lvs is not None:
            rhs = self.parse_expr()
            quiet = self.end_of_statement()
            return ("assign", lvs, rhs, quiet)
        self.p = save
        e = self.parse_expr()
        quiet = self.end_of_statement()
        return ("expr", e, quiet)

    def try_parse_lhs(self):
        """[a, b.c, d(i)] = ...  or  a(i).f = ... ; returns list of l-values or None."""
        try:
            if self.is_op("["):
                # find matching ] and check that '=' (not '==') follows
                depth, k = 0, 0
                while True:
                    tok = self.peek(k)
                    if tok.kind == "eof" or tok.kind == "nl":
                        return None
                    if tok.kind == "op" and tok.val in "([{":
                        depth += 1
                    elif tok.kind == "op" and tok.val in ")]}":
                        depth -= 1
                        if depth == 0:
                            break
                    k += 1
                if not self.is_op("=", k + 1):
                    return None
                self.next()
                lvs = []
                while not self.is_op("]"):
                    if self.is_op(","):
                        self.next()
                        continue
                    if self.is_op("~"):
                        self.next()
                        lvs.append(None)
                        continue
                    lvs.append(self.parse_lvalue())
                self.next()
                self.expect_op("=")
                return lvs
            if self.peek().kind != "id":
                return None
            lv = self.parse_lvalue()
            if self.is_op("=") :
                self.next()
                return [lv]
            return None
        except MSyntaxError:
            return None

    def parse_lvalue(self):
        tok = self.next()
        if tok.kind != "id":
            self.err("l-value expected", tok)
        acc = self.parse_accessors(in_matrix=False)
        return ("ref", tok.val, acc)

    def parse_if(self):
        self.next()
        clauses = []
        cond = self.parse_expr()
        body = self.parse_block(("elseif", "else", "end"))
        clauses.append((cond, body))
        else_body = None
        while True:
            if self.is_kw("elseif"):
                self.next()
                cond = self.parse_expr()
                body = self.parse_block(("elseif", "else", "end"))
                clauses.append((cond, body))
            elif self.is_kw("else"):
                self.next()
                else_body = self.parse_block(("end",))
            else:
                break
        if not self.is_kw("end"):
            self.err("`end` expected to close `if`")
        self.next()
        return ("if", clauses, else_body)

    def parse_for(self):
        self.next()
        paren = False
        if self.is_op("("):
            paren = True
            self.next()
        var = self.next().val
        self.expect_op("=")
        e = self.parse_expr()
        if paren:
            self.expect_op(")")
        body = self.parse_block(("end",))
        self.next()
        return ("for", var, e, body)

    # -- expressions ------------------------------------------------------------------------
    # in_matrix: we are directly inside [ ] or { } where whitespace separates elements
    def parse_expr(self, in_matrix=False):
        return self.parse_oror(in_matrix)

    def _binary_level(self, ops, sub, in_matrix):
        left = sub(in_matrix)
        while True:
            tok = self.peek()
            if tok.kind != "op" or tok.val not in ops:
                return left
            if in_matrix and tok.val in ("+", "-") and tok.ws and not self.peek(1).ws:
                return left       # [a -b] : unary sign of a new element
            self.next()
            right = sub(in_matrix)
            left = ("bin", tok.val, left, right)

    def parse_oror(self, m):
        return self._binary_level(("||",), self.parse_andand, m)

    def parse_andand(self, m):
        return self._binary_level(("&&",), self.parse_or, m)

    def parse_or(self, m):
        return self._binary_level(("|",), self.parse_and, m)

    def parse_and(self, m):
        return self._binary_level(("&",), self.parse_cmp, m)

    def parse_cmp(self, m):
        return self._binary_level(("==", "~=", "<", "<=", ">", ">="), self.parse_range, m)

    def parse_range(self, m):
        first = self.parse_additive(m)
        if self.is_op(":") and not self._colon_is_index_all():
            self.next()
            second = self.parse_additive(m)
            if self.is_op(":") and not self._colon_is_index_all():
                self.next()
                third = self.parse_additive(m)
                return ("range", first, second, third)
            return ("range", first, None, second)
        return first

    def _colon_is_index_all(self):
        nxt = self.peek(1)
        return nxt.kind == "op" and nxt.val in (")", ",")

    def parse_additive(self, m):
        return self._binary_level(("+", "-"), self.parse_mul, m)

    def parse_mul(self, m):
        return self._binary_level(("*", "/", "\\", ".*", "./", ".\\"), self.parse_unary, m)

    def parse_unary(self, m):
        tok = self.peek()
        if tok.kind == "op" and tok.val in ("+", "-", "~", "!"):
            self.next()
            operand = self.parse_unary(m)
            return ("un", "~" if tok.val == "!" else tok.val, operand)
        return self.parse_power(m)

    def parse_power(self, m):
        base = self.parse_postfix(m)
        while self.is_op("^") or self.is_op(".^"):
            op = self.next().val
            # the exponent may carry its own unary sign: 2^-1
            tok = self.peek()
            if tok.kind == "op" and tok.val in ("+", "-", "~"):
                self.next()
                expo = ("un", tok.val, self.parse_postfix(m))
            else:
                expo = self.parse_postfix(m)
            base = ("bin", op, base, expo)
        return base

    def parse_postfix(self, m):
        e = self.parse_primary(m)
        while True:
            tok = self.peek()
            if tok.kind == "op" and tok.val in ("'", ".'"):
                if m and tok.ws:
                    return e
                self.next()
                e = ("post", "'", e)
                continue
            return e

    def parse_accessors(self, in_matrix):
        acc = []
        while True:
            tok = self.peek()
            if tok.kind != "op":
                break
            if tok.val == "(":
                if in_matrix and tok.ws:
                    break                  # [a (1)] : two elements
                self.next()
                acc.append(("()", self.parse_args(")")))
            elif tok.val == "{":
                if in_matrix and tok.ws:
                    break
                self.next()
                acc.append(("{}", self.parse_args("}")))
            elif tok.val == "." and self.peek(1).kind in ("id", "kw") and not self.peek(1).ws:
                self.next()
                acc.append((".", self.next().val))
            else:
                break
        return acc

    def parse_args(self, closer):
        args = []
        while True:
            while self.peek().kind == "nl":
                self.next()
            if self.is_op(closer):
                self.next()
                return args
            if self.is_op(","):
                self.next()
                continue
            if self.is_op(":") and self._colon_is_index_all_here(closer):
                self.next()
                args.append(("colon",))
                continue
            args.append(self.parse_expr(False))

    def _colon_is_index_all_here(self, closer):
        nxt = self.peek(1)
        return nxt.kind == "op" and nxt.val in (closer, ",")

    def parse_primary(self, m):
        tok = self.next()
        if tok.kind == "num":
            return ("num", tok.val)
        if tok.kind == "str":
            return ("str", tok.val)
        if tok.kind == "id":
            acc = self.parse_accessors(m)
            return ("ref", tok.val, acc)
        if tok.kind == "op":
            if tok.val == "(":
                e = self.parse_expr(False)
                self.expect_op(")")
                # a parenthesised expression can still be indexed in Octave, not in MATLAB: not supported
                return ("paren", e)
            if tok.val == "[":
                return ("matrix", self.parse_rows("]"))
            if tok.val == "{":
                return ("cell", self.parse_rows("}"))
            if tok.val == "end":
                return ("end",)
            if tok.val == ":":
                return ("colon",)
        self.err("unexpected token %r" % (tok.val,), tok)

    def parse_rows(self, closer):
        rows, row = [], []
        while True:
            tok = self.peek()
            if tok.kind == "op" and tok.val == closer:
                self.next()
                if row:
                    rows.append(row)
                return rows
            if tok.kind == "nl" or (tok.kind == "op" and tok.val == ";"):
                self.next()
                if row:
                    rows.append(row)
                    row = []
                continue
            if tok.kind == "op" and tok.val == ",":
                self.next()
                continue
            if tok.kind == "eof":
                self.err("unterminated matrix literal")
            row.append(self.parse_expr(True))


def parse_source(src, fname="<m>"):
    return Parser(src, fname).parse_file()
