"""A tiny interpreter for the INTEGER subset of Fortran 77 the reference's index / descriptor tools are written in
(TOOLS/numroc.f, indxg2p.f, indxg2l.f, indxl2g.f, iceil.f, ilcm.f, infog2l.f, descinit.f, chk1mat.f): fixed form, INTEGER scalars and
1-based arrays, PARAMETER, assignments, block IF / ELSE IF / ELSE / END IF, logical IF, GO TO, labelled CONTINUE, CALL, RETURN, the
intrinsics MOD / MAX / MIN / ABS, DO loops, DOUBLE PRECISION scalars, LSAME, and calls to other units or to Python callbacks
(the latter is how SRC/pdgetrf.f, pdgetf2.f, pdlaswp.f and pdgetrs.f are run with numpy standing in for the PBLAS leaves).

TEST INFRASTRUCTURE.  There is no Fortran compiler in this image, so this is how the tests execute the reference's OWN source text of
those routines (read from /root/reference at test time, never copied) and pin the product's and the oracle's restatements against it
instead of against each other.  tests/golden/make_tools_golden.py stores the values it produces so the check also travels."""
import re

_TOKEN = re.compile(r"\s*(?:(\d+\.(?![A-Z]+\.)\d*(?:[DE][-+]?\d+)?|\d+[DE][-+]?\d+|\d+)|('(?:[^']|'')*')|(\.[A-Z]+\.)|([A-Z_][A-Z0-9_]*)|(\*\*|[-+*/(),=]))")


def _idiv(a, b):
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q          # Fortran integer division truncates toward zero


def _imod(a, b):
    return a - _idiv(a, b) * b                          # MOD has the sign of the first argument


class Unit:
    def __init__(self, kind, name, args, stmts):
        self.kind, self.name, self.args, self.stmts = kind, name, args, stmts
        self.labels = {lab: i for i, (lab, _) in enumerate(stmts) if lab}
        decl = " ".join(st for _, st in stmts if re.match(r"(INTEGER|DOUBLE|REAL|COMPLEX|LOGICAL|CHARACTER)\b", st) and "=" not in st.split("(")[0])
        self.array_args = {a for a in args if re.search(r"\b%s\s*\(" % re.escape(a), decl)}      # dummies declared with dimensions
        self.saves_all = any(st == "SAVE" for _, st in stmts)                                    # a bare SAVE: every local persists
        # leading dimension of the arrays declared with two dimensions, e.g. A( LDA, * ): needed when the actual is a flat 1-D view
        self.lead = {m.group(1): m.group(2).strip() for m in re.finditer(r"\b([A-Z_][A-Z0-9_]*)\s*\(([^(),]+),[^()]*\)", decl)}


def parse(text):
    """-> Unit.  One program unit per file (the reference's TOOLS files)."""
    lines = []
    for raw in text.splitlines():
        if not raw.strip() or raw[0] in "*Cc!":
            continue
        raw = raw.rstrip("\n").upper().split("!")[0] if "'" not in raw else raw.rstrip("\n")
        body = raw[6:72] if len(raw) > 6 else ""
        if len(raw) > 5 and raw[5] not in " 0":          # continuation line
            lines[-1][1] += " " + body.strip()
        else:
            lines.append([raw[:5].strip(), body.strip()])
    up = lambda s: re.sub(r"'[^']*'|[^']+", lambda m: m.group(0) if m.group(0).startswith("'") else m.group(0).upper(), s)
    lines = [(lab, up(st)) for lab, st in lines]
    head = lines[0][1]
    m = re.match(r"(?:(INTEGER|REAL|DOUBLE\s+PRECISION|LOGICAL)\s+)?(FUNCTION|SUBROUTINE)\s+([A-Z0-9_]+)\s*\((.*)\)\s*$", head)
    assert m, head
    args = [a.strip() for a in m.group(4).split(",") if a.strip()]
    return Unit(m.group(2), m.group(3), args, lines[1:])


def parse_file(text):
    """-> [Unit, ...] for a file that holds several program units (each ends with a line END)"""
    units, cur = [], []
    for raw in text.splitlines():
        cur.append(raw)
        if raw[:1] not in "*Cc!" and raw[6:72].strip().upper() == "END":
            if any(l.strip() and l[:1] not in "*Cc!" for l in cur[:-1]):
                units.append(parse("\n".join(cur)))
            cur = []
    return units


def _wrap32(v):
    """INTEGER arithmetic of the reference is 32-bit two's complement (the generator's LMUL overflows and corrects for it, pmatgeninc.f)"""
    return ((v + 2 ** 31) % 2 ** 32) - 2 ** 31


class Interp:
    def __init__(self, units=(), callbacks=None):
        self.units = {u.name: u for u in units}
        self.callbacks = dict(callbacks or {})          # NAME -> f(interp, env, arg_exprs) for CALLs; NAME -> f(*values) for functions
        self.common = {}                                 # COMMON block name -> {variable name: storage}
        self.saved = {}                                  # unit name -> locals kept between calls (units with a bare SAVE)
        self.wrap32 = False                              # wrap INTEGER + - * to 32 bits

    # ---- expressions --------------------------------------------------------------------------------------------------------
    def eval(self, text, env):
        toks = []
        pos = 0
        while pos < len(text):
            m = _TOKEN.match(text, pos)
            if not m:
                if text[pos:].strip() == "":
                    break
                raise SyntaxError(text[pos:])
            toks.append(m.group(m.lastindex))
            pos = m.end()
        self._t, self._i, self._env = toks, 0, env
        v = self._or()
        assert self._i == len(toks), (text, toks[self._i:])
        return v

    def _peek(self):
        return self._t[self._i] if self._i < len(self._t) else None

    def _next(self):
        self._i += 1
        return self._t[self._i - 1]

    def _or(self):
        v = self._and()
        while self._peek() == ".OR.":
            self._next(); w = self._and(); v = bool(v) or bool(w)
        return v

    def _and(self):
        v = self._not()
        while self._peek() == ".AND.":
            self._next(); w = self._not(); v = bool(v) and bool(w)
        return v

    def _not(self):
        if self._peek() == ".NOT.":
            self._next()
            return not self._not()
        return self._rel()

    def _rel(self):
        v = self._add()
        ops = {".LT.": lambda a, b: a < b, ".LE.": lambda a, b: a <= b, ".EQ.": lambda a, b: a == b, ".NE.": lambda a, b: a != b,
               ".GT.": lambda a, b: a > b, ".GE.": lambda a, b: a >= b}
        if self._peek() in ops:
            f = ops[self._next()]
            v = f(v, self._add())
        return v

    def _add(self):
        if self._peek() == "-":
            self._next(); v = -self._mul()
        elif self._peek() == "+":
            self._next(); v = self._mul()
        else:
            v = self._mul()
        while self._peek() in ("+", "-"):
            op = self._next(); w = self._mul()
            v = v + w if op == "+" else v - w
            if self.wrap32 and isinstance(v, int) and not isinstance(v, bool):
                v = _wrap32(v)
        return v

    def _mul(self):
        v = self._pow()
        while self._peek() in ("*", "/"):
            op = self._next(); w = self._pow()
            v = v * w if op == "*" else (_idiv(v, w) if isinstance(v, int) and isinstance(w, int) else v / w)
            if self.wrap32 and isinstance(v, int) and not isinstance(v, bool):
                v = _wrap32(v)
        return v

    def _pow(self):
        v = self._prim()
        if self._peek() == "**":
            self._next(); v = v ** self._pow()
        return v

    def _prim(self):
        t = self._next()
        if t == "(":
            v = self._or(); assert self._next() == ")"
            return v
        if t == "-":
            return -self._prim()
        if t.isdigit():
            return int(t)
        if t[0].isdigit():
            return float(t.replace("D", "E"))
        if t == ".TRUE.":
            return True
        if t == ".FALSE.":
            return False
        if t.startswith("'"):
            return t[1:-1]
        name = t
        if self._peek() == "(" and name in getattr(self, "raw_functions", {}):
            # a function that needs its arguments BY ADDRESS (e.g. DDOT( N, A( IOFFA ), 1, ... )): hand over the argument texts
            self._next()
            parts, depth, cur = [], 0, []
            while True:
                tk = self._next()
                if tk == "(":
                    depth += 1
                elif tk == ")":
                    if depth == 0:
                        break
                    depth -= 1
                if tk == "," and depth == 0:
                    parts.append(" ".join(cur)); cur = []
                else:
                    cur.append(tk)
            parts.append(" ".join(cur))
            saved = (self._t, self._i, self._env)
            v = self.raw_functions[name](self, self._env, parts)
            self._t, self._i, self._env = saved
            return v
        if self._peek() == "(":
            self._next()
            args = []
            if self._peek() != ")":
                args.append(self._or())
                while self._peek() == ",":
                    self._next(); args.append(self._or())
            assert self._next() == ")"
            env = self._env
            if name in env and hasattr(env[name], "__getitem__") and not isinstance(env[name], str):
                if len(args) == 2 and getattr(env[name], "ndim", 2) == 1:
                    v2 = env[name][args[0] - 1 + (args[1] - 1) * self._lead(name, env)]
                    return complex(v2) if getattr(v2, "dtype", None) == "complex128" else float(v2)
                if len(args) == 2:
                    v2 = env[name][args[0] - 1, args[1] - 1]
                    return complex(v2) if isinstance(v2, complex) or getattr(v2, "dtype", None) == "complex128" else float(v2)
                return env[name][args[0] - 1]
            if name == "LSAME":
                return str(args[0])[:1].upper() == str(args[1])[:1].upper()
            if name in ("DBLE", "REAL"):
                return float(args[0].real) if isinstance(args[0], complex) else float(args[0])
            if name == "DCMPLX":
                return complex(args[0], args[1] if len(args) > 1 else 0.0)
            if name == "DCONJG":
                return complex(args[0]).conjugate()
            if name == "DIMAG":
                return float(complex(args[0]).imag)
            if name == "ICHAR":
                return ord(str(args[0])[:1])
            if name == "SQRT":
                return float(args[0]) ** 0.5
            if name == "SIGN":
                return abs(args[0]) if str(args[1])[0] != "-" and args[1] >= 0 else -abs(args[0])
            if name == "NINT":
                return int(round(args[0]))
            if name == "MOD":
                return _imod(*args)
            if name == "MAX":
                return max(args)
            if name == "MIN":
                return min(args)
            if name == "ABS":
                return abs(args[0])
            if name in self.units:
                saved = (self._t, self._i, self._env)
                out = self.call(name, *args)
                self._t, self._i, self._env = saved
                return out["__result__"]
            return self.callbacks[name](*args)
        return self._env[name]

    def address(self, part, env):
        """('NAME( expr )' or 'NAME') -> (array object, 0-based offset): Fortran's pass-by-address of an array element"""
        m = re.match(r"\s*([A-Z_][A-Z0-9_]*)\s*(?:\((.*)\))?\s*$", part)
        assert m, part
        return env[m.group(1)], (self.eval(m.group(2), env) - 1 if m.group(2) else 0)

    def assign(self, target, env, value):
        """target: a variable name or an array element as written in the source, e.g. 'IPIV( IIA+J-JA )'"""
        m = re.match(r"\s*([A-Z_][A-Z0-9_]*)\s*(\((.*)\))?\s*$", target)
        assert m, target
        if m.group(2):
            self._store(env, m.group(1), m.group(3), value)
        else:
            env[m.group(1)] = value

    def _store(self, env, name, index_text, value):
        parts, depth, cur = [], 0, ""
        for ch in index_text:
            if ch == "," and depth == 0:
                parts.append(cur); cur = ""
            else:
                depth += ch == "("; depth -= ch == ")"; cur += ch
        parts.append(cur)
        idx = [self.eval(p_, env) - 1 for p_ in parts]
        if len(idx) == 2 and getattr(env[name], "ndim", 2) == 1:
            env[name][idx[0] + idx[1] * self._lead(name, env)] = value
        elif len(idx) == 2:
            env[name][idx[0], idx[1]] = value
        else:
            env[name][idx[0]] = value

    def _lead(self, name, env):
        """value of the declared leading dimension of the two-dimensional array `name` in the unit being executed"""
        saved = (self._t, self._i, self._env)
        v = self.eval(self._unit_stack[-1].lead[name], env)
        self._t, self._i, self._env = saved
        return v

    # ---- statements ---------------------------------------------------------------------------------------------------------
    def call(self, name, *args):
        """Runs unit `name`; scalars by value (their final values come back in the result dict), arrays as Python lists (in place)."""
        if not hasattr(self, "_unit_stack"):
            self._unit_stack = []
        self._unit_stack.append(self.units[name])
        try:
            return self._call(name, *args)
        finally:
            self._unit_stack.pop()

    def _call(self, name, *args):
        u = self.units[name]
        env = dict(zip(u.args, args))
        if u.kind == "FUNCTION":
            env[u.name] = 0
        st = u.stmts
        pc = 0

        def skip_to_branch(i):
            """from a false IF / ELSE IF at i: index of the next ELSE IF / ELSE / END IF of the same block"""
            depth = 0
            j = i + 1
            while True:
                s = st[j][1]
                if re.match(r"IF\s*\(.*\)\s*THEN$", s):
                    depth += 1
                elif re.match(r"END\s*IF$", s):
                    if depth == 0:
                        return j
                    depth -= 1
                elif depth == 0 and (re.match(r"ELSE\s*IF\s*\(.*\)\s*THEN$", s) or s == "ELSE"):
                    return j
                j += 1

        def skip_to_end(i):
            depth = 0
            j = i + 1
            while True:
                s = st[j][1]
                if re.match(r"IF\s*\(.*\)\s*THEN$", s):
                    depth += 1
                elif re.match(r"END\s*IF$", s):
                    if depth == 0:
                        return j
                    depth -= 1
                j += 1

        def split_cond(s):
            """'IF ( cond ) rest' -> (cond, rest)"""
            i = s.index("(")
            depth, j = 0, i
            while True:
                if s[j] == "(":
                    depth += 1
                elif s[j] == ")":
                    depth -= 1
                    if depth == 0:
                        break
                j += 1
            return s[i + 1:j], s[j + 1:].strip()

        def simple(s):
            """executes a non-block statement; returns 'RETURN', ('GOTO', label) or None"""
            if s in ("CONTINUE",):
                return None
            if s in ("RETURN", "END"):
                return "RETURN"
            m = re.match(r"GO\s*TO\s*(\d+)$", s)
            if m:
                return ("GOTO", m.group(1))
            m = re.match(r"GO\s*TO\s*\(([\d,\s]+)\)\s*,?\s*(.+)$", s)
            if m:                                         # computed GO TO: the k-th label, or fall through when k is out of range
                labs, k_ = [x.strip() for x in m.group(1).split(",")], self.eval(m.group(2), env)
                return ("GOTO", labs[k_ - 1]) if 1 <= k_ <= len(labs) else None
            m = re.match(r"CALL\s+([A-Z0-9_]+)\s*\((.*)\)$", s)
            if m:
                cname, inner = m.group(1), m.group(2)
                parts, depth, cur = [], 0, ""
                for ch in inner:
                    if ch == "," and depth == 0:
                        parts.append(cur.strip()); cur = ""
                    else:
                        depth += ch == "("; depth -= ch == ")"; cur += ch
                parts.append(cur.strip())
                if cname in self.units:                  # another interpreted unit: by reference = copy in, copy back
                    callee = self.units[cname]
                    # an actual argument that is a not-yet-defined variable is an output of the callee: pass a placeholder
                    vals = []
                    for p_, d_ in zip(parts, callee.args):
                        m_ = re.fullmatch(r"([A-Z_][A-Z0-9_]*)\s*\((.*)\)", p_)
                        if d_ in callee.array_args and m_ and m_.group(1) in env and hasattr(env[m_.group(1)], "shape"):
                            vals.append(env[m_.group(1)][self.eval(m_.group(2), env) - 1:])      # numpy view: Fortran's pass-by-address
                        else:
                            vals.append(env[p_] if p_ in env else (0 if re.fullmatch(r"[A-Z_][A-Z0-9_]*", p_) else self.eval(p_, env)))
                    out = self.call(cname, *vals)
                    for p_, d_ in zip(parts, callee.args):
                        v_ = out[d_]
                        if isinstance(v_, (int, float, bool, str)):
                            m_ = re.fullmatch(r"([A-Z_][A-Z0-9_]*)\s*(\(.*\))?", p_)
                            if m_ and m_.group(2):               # an array element only if ITS parenthesis closes at the very end
                                depth_ = 0
                                for k_, ch_ in enumerate(m_.group(2)):
                                    depth_ += ch_ == "("; depth_ -= ch_ == ")"
                                    if depth_ == 0 and k_ < len(m_.group(2)) - 1:
                                        m_ = None
                                        break
                            if m_ and (m_.group(2) is None or (m_.group(1) in env and hasattr(env[m_.group(1)], "__setitem__"))):
                                self.assign(p_, env, v_)
                    return None
                self.callbacks[cname](self, env, parts)
                return None
            depth, eq = 0, -1
            for k_, ch_ in enumerate(s):
                depth += ch_ == "("; depth -= ch_ == ")"
                if ch_ == "=" and depth == 0:
                    eq = k_; break
            assert eq > 0, s
            m = re.match(r"([A-Z_][A-Z0-9_]*)\s*(\(.*\))?\s*$", s[:eq])
            assert m, s
            m = type("M", (), {"group": lambda self_, i_, m_=m, rhs_=s[eq + 1:].strip(): rhs_ if i_ == 3 else m_.group(i_)})()
            val = self.eval(m.group(3), env)
            if m.group(2):
                if m.group(1) not in env:
                    env[m.group(1)] = [0] * 64             # a small local work array (e.g. IDUM1( 1 ), DESCIP( DLEN_ ))
                self._store(env, m.group(1), m.group(2)[1:-1], val)
            else:
                env[m.group(1)] = val
            return None

        loops = []                                       # active DO loops: [label, var, last, step, body_start]
        while pc < len(st):
            s = st[pc][1]
            if re.match(r"(IMPLICIT|INTEGER|EXTERNAL|INTRINSIC|LOGICAL|CHARACTER|DOUBLE|REAL|COMPLEX)\b", s) and "=" not in s.split("(")[0]:
                if re.match(r"(INTEGER|DOUBLE)\b", s):       # local arrays such as IDUM1( 1 ), DESCIP( DLEN_ ): small work arrays
                    for nm in re.findall(r"([A-Z_][A-Z0-9_]*)\s*\(", s):
                        if nm not in env and nm not in ("PRECISION",):
                            env[nm] = [0] * 64
                pc += 1; continue
            if s.startswith("SAVE"):
                if s == "SAVE":
                    env.update(self.saved.get(u.name, {}))
                pc += 1; continue
            m = re.match(r"COMMON\s*/\s*([A-Z0-9_]+)\s*/\s*(.*)$", s)
            if m:
                blk = self.common.setdefault(m.group(1), {})
                for nm in [x.strip() for x in m.group(2).split(",")]:
                    env[nm] = blk.setdefault(nm, [0] * 8)  # the common blocks met here hold small integer arrays
                pc += 1; continue
            m = re.match(r"DO\s+(\d+)\s+([A-Z_][A-Z0-9_]*)\s*=\s*(.*)$", s)
            if m:
                bounds, depth, cur = [], 0, ""
                for ch in m.group(3):
                    if ch == "," and depth == 0:
                        bounds.append(cur); cur = ""
                    else:
                        depth += ch == "("; depth -= ch == ")"; cur += ch
                bounds.append(cur)
                first, last = self.eval(bounds[0], env), self.eval(bounds[1], env)
                step = self.eval(bounds[2], env) if len(bounds) > 2 else 1
                env[m.group(2)] = first
                if (step > 0 and first > last) or (step < 0 and first < last):
                    pc = u.labels[m.group(1)] + 1         # zero-trip loop
                else:
                    loops.append([m.group(1), m.group(2), last, step, pc + 1]); pc += 1
                continue
            if s.startswith("PARAMETER"):
                for part in re.findall(r"([A-Z_][A-Z0-9_]*)\s*=\s*([^,()]+(?:\([^()]*\))?[^,()]*)", s[s.index("(") + 1:s.rindex(")")]):
                    env[part[0]] = self.eval(part[1].strip(), env)
                pc += 1; continue
            m = re.match(r"(ELSE\s*)?IF\s*\(", s)
            if m and s.endswith("THEN"):
                if m.group(1):                           # ELSE IF reached by falling out of a taken branch
                    pc = skip_to_end(pc); continue
                cond, _ = split_cond(s)
                while True:
                    if self.eval(cond, env):
                        pc += 1; break
                    pc = skip_to_branch(pc)
                    s2 = st[pc][1]
                    if s2 == "ELSE":
                        pc += 1; break
                    if re.match(r"END\s*IF$", s2):
                        pc += 1; break
                    cond, _ = split_cond(s2[s2.index("IF"):])
                continue
            if s == "ELSE":
                pc = skip_to_end(pc); continue
            if re.match(r"END\s*IF$", s):
                pc += 1; continue
            if re.match(r"IF\s*\(", s):                  # logical IF
                cond, rest = split_cond(s)
                r = simple(rest) if self.eval(cond, env) else None
            else:
                r = simple(s)
            if r == "RETURN":
                break
            if isinstance(r, tuple):
                pc = u.labels[r[1]]
                while loops and not (loops[-1][4] <= pc <= u.labels[loops[-1][0]]):
                    loops.pop()                          # a jump out of (nested) DO loops ends them (pdmatgen.f: GO TO 270 / 300)
                continue
            if loops and st[pc][0] == loops[-1][0]:      # the terminal statement of the innermost DO loop
                lab, var, last, step, start = loops[-1]
                env[var] += step
                if (step > 0 and env[var] <= last) or (step < 0 and env[var] >= last):
                    pc = start; continue
                loops.pop()
            pc += 1
        if u.saves_all:
            self.saved[u.name] = {k: v for k, v in env.items() if k not in u.args}
        out = {k: env[k] for k in u.args}
        if u.kind == "FUNCTION":
            out["__result__"] = env[u.name]
        return out


def load_tools(ref_root, grid=None, errors=None):
    """The reference's TOOLS units with BLACS_GRIDINFO answering from `grid` = (nprow, npcol, myrow, mycol) and PXERBLA recorded."""
    import os
    units = [parse(open(os.path.join(ref_root, "TOOLS", f + ".f")).read())
             for f in ("numroc", "indxg2p", "indxg2l", "indxl2g", "iceil", "ilcm", "infog2l", "descinit", "chk1mat")]
    state = {"grid": grid or (1, 1, 0, 0)}

    def gridinfo(interp, env, parts):
        for name, v in zip(parts[1:], state["grid"]):
            env[name] = v

    def pxerbla(interp, env, parts):
        if errors is not None:
            errors.append((interp.eval(parts[1], env), interp.eval(parts[2], env)))
    it = Interp(units, {"BLACS_GRIDINFO": gridinfo, "PXERBLA": pxerbla})
    it.state = state
    return it
