#!/usr/bin/env python
"""Tiny fixed-form FORTRAN-77 subset -> C transpiler.  TEST INFRASTRUCTURE ONLY.

Why: the reference's native kernels (solvshift/src/{efprot,solpol2,shftex,exrep}.f) are F77 and
this image has no Fortran compiler.  This script reads those sources WHERE THEY LIE under
/root/reference and writes C translations into oracle/_ref/ (git-ignored; never committed), which
``oracle/build_ref.py`` compiles with gcc into oracle/_ref/libslvref.so.  That library is the
reference's own arithmetic and is used to validate the NumPy restatement in
oracle/solefp_oracle.py and as the CPU baseline of the native kernels.

Supported subset (exactly what those four files use): SUBROUTINE, IMPLICIT DOUBLE PRECISION(A-H,O-Z),
DIMENSION / DOUBLE PRECISION / INTEGER / LOGICAL / CHARACTER declarations with adjustable arrays,
PARAMETER, DATA (zero-fill only), COMMON, EXTERNAL, labelled and block DO, block and logical IF,
GOTO, CONTINUE, CALL, RETURN, assignments, ``**``, .EQ./.NE./.LT./.LE./.GT./.GE./.AND./.OR./.NOT.,
intrinsics DSQRT DABS DLOG DEXP DFLOAT.  WRITE/OPEN/CLOSE are dropped (diagnostics only).
Calling convention: every argument is a pointer, symbols are lower-case with a trailing underscore.
"""
import re
import sys

INTRINSICS = {'DSQRT': ('sqrt', 'd'), 'DABS': ('fabs', 'd'), 'DLOG': ('log', 'd'), 'DEXP': ('exp', 'd'),
              'DFLOAT': ('(double)', 'd'), 'SQRT': ('sqrt', 'd'), 'ABS': ('fabs', 'd'), 'DBLE': ('(double)', 'd')}
EXTERNAL_FUNCS = {'DDOT': 'd', 'ILAENV': 'i'}
SKIP_UNITS = {'MATWRT', 'VECWRT'}          # file I/O helpers with CHARACTER(*) arguments
SKIP_CALLS = {'MATWRT', 'VECWRT'}
EXTERNAL_SUBS = {'DGEMM', 'DGETRF', 'DGETRI'}


def logical_lines(text):
    """Join continuation lines, drop comments; yields (label, statement)."""
    out = []
    for raw in text.split('\n'):
        if not raw.strip():
            continue
        if raw[0] in 'Cc*!':
            continue
        line = raw.rstrip('\n')
        if '\t' in line[:6]:
            line = line.replace('\t', '      ', 1)
        if len(line) > 5 and line[5] not in ' 0' and line[:5].strip() == '':
            out[-1][1] += line[6:72]
            continue
        label = line[:5].strip()
        out.append([label, line[6:72]])
    res = []
    for label, st in out:
        # strip inline '!' comments outside strings
        s = ''; inq = False
        for ch in st:
            if ch in "'\"":
                inq = not inq
            if ch == '!' and not inq:
                break
            s += ch
        res.append((label, s.strip()))
    return res


TOK = re.compile(r"\s*(?:(\d+\.?\d*(?:[DEde][+-]?\d+)?|\.\d+(?:[DEde][+-]?\d+)?)|(\.[A-Za-z]+\.)|([A-Za-z_][A-Za-z0-9_]*)|('(?:[^']|'')*'|\"[^\"]*\")|(\*\*|[-+*/(),=:]))")


def tokenize(s):
    toks = []; pos = 0
    s = s.strip()
    while pos < len(s):
        m = TOK.match(s, pos)
        if not m:
            raise SyntaxError('cannot tokenise: %r at %r' % (s, s[pos:]))
        num, dot, ident, string, op = m.groups()
        if num is not None:
            # "1.EQ." ambiguity: a number followed by .EQ. -- regex is greedy on '1.'; repair
            if num.endswith('.') and re.match(r'[A-Za-z]+\.', s[m.end():]):
                toks.append(('num', num[:-1])); pos = m.end() - 1; continue
            toks.append(('num', num))
        elif dot is not None:
            toks.append(('dot', dot.upper()))
        elif ident is not None:
            toks.append(('id', ident.upper()))
        elif string is not None:
            toks.append(('str', string[1:-1]))
        else:
            toks.append(('op', op))
        pos = m.end()
    return toks


class Unit(object):
    def __init__(self, name, args):
        self.name = name; self.args = args
        self.types = {}        # name -> 'd'|'i'|'l'|'c'
        self.dims = {}         # name -> list of dim strings (token lists)
        self.params = {}       # name -> expr tokens
        self.common = {}       # name -> global C name
        self.body = []         # (label, statement)
        self.implicit_double = False
        self.data_zero = set()

    def typeof(self, name):
        if name in self.types:
            return self.types[name]
        if name[0] in 'IJKLMN':
            return 'i'
        return 'd'


class ExprParser(object):
    def __init__(self, unit, toks):
        self.u = unit; self.t = toks; self.p = 0

    def peek(self):
        return self.t[self.p] if self.p < len(self.t) else (None, None)

    def next(self):
        tok = self.peek(); self.p += 1; return tok

    def expect(self, val):
        k, v = self.next()
        if v != val:
            raise SyntaxError('expected %r got %r in %r' % (val, v, self.t))

    def parse(self):
        r = self.p_or()
        return r

    def p_or(self):
        c, t = self.p_and()
        while self.peek() == ('dot', '.OR.'):
            self.next(); c2, _ = self.p_and(); c = '(%s || %s)' % (c, c2); t = 'l'
        return c, t

    def p_and(self):
        c, t = self.p_not()
        while self.peek() == ('dot', '.AND.'):
            self.next(); c2, _ = self.p_not(); c = '(%s && %s)' % (c, c2); t = 'l'
        return c, t

    def p_not(self):
        if self.peek() == ('dot', '.NOT.'):
            self.next(); c, _ = self.p_rel(); return '(!%s)' % c, 'l'
        return self.p_rel()

    RELOPS = {'.EQ.': '==', '.NE.': '!=', '.LT.': '<', '.LE.': '<=', '.GT.': '>', '.GE.': '>='}

    def p_rel(self):
        c, t = self.p_arith()
        k, v = self.peek()
        if k == 'dot' and v in self.RELOPS:
            self.next(); c2, t2 = self.p_arith()
            if t == 'c' or t2 == 'c':
                # character comparison: compare first characters
                def ch(x, tx):
                    return x if tx == 'c1' else x
                c = '(%s == %s)' % (c, c2) if v == '.EQ.' else '(%s != %s)' % (c, c2)
            else:
                c = '(%s %s %s)' % (c, self.RELOPS[v], c2)
            t = 'l'
        elif k == 'dot' and v in ('.TRUE.', '.FALSE.'):
            pass
        return c, t

    def p_arith(self):
        k, v = self.peek()
        if (k, v) == ('op', '-'):
            self.next(); c, t = self.p_term(); c = '(-%s)' % c
        elif (k, v) == ('op', '+'):
            self.next(); c, t = self.p_term()
        else:
            c, t = self.p_term()
        while self.peek() in (('op', '+'), ('op', '-')):
            _, op = self.next(); c2, t2 = self.p_term()
            c = '(%s %s %s)' % (c, op, c2); t = 'd' if 'd' in (t, t2) else t
        return c, t

    def p_term(self):
        c, t = self.p_power()
        while self.peek() in (('op', '*'), ('op', '/')):
            _, op = self.next(); c2, t2 = self.p_power()
            c = '(%s %s %s)' % (c, op, c2); t = 'd' if 'd' in (t, t2) else t
        return c, t

    def p_power(self):
        c, t = self.p_primary()
        if self.peek() == ('op', '**'):
            self.next()
            k, v = self.peek()
            if (k, v) == ('op', '-'):
                self.next(); c2, t2 = self.p_power(); c2 = '(-%s)' % c2
            else:
                c2, t2 = self.p_power()
            if t2 == 'i' and c2.strip('()') == '2':
                c = '((%s)*(%s))' % (c, c)
            else:
                c = 'pow((double)(%s),(double)(%s))' % (c, c2); t = 'd'
        return c, t

    def p_args(self):
        args = []
        self.expect('(')
        if self.peek() == ('op', ')'):
            self.next(); return args
        while True:
            start = self.p
            c, t = self.p_or()
            args.append((c, t, self.t[start:self.p]))
            k, v = self.next()
            if v == ')':
                break
            if v != ',':
                raise SyntaxError('bad arg list %r' % (self.t,))
        return args

    def p_primary(self):
        k, v = self.next()
        if k == 'num':
            if re.search(r'[.DEde]', v):
                return v.upper().replace('D', 'e').replace('E', 'e'), 'd'
            return v, 'i'
        if k == 'str':
            if len(v) == 1:
                return "'%s'" % v, 'c'
            return '"%s"' % v, 'c'
        if k == 'dot':
            if v == '.TRUE.':
                return '1', 'l'
            if v == '.FALSE.':
                return '0', 'l'
            raise SyntaxError('unexpected %s' % v)
        if (k, v) == ('op', '('):
            c, t = self.p_or(); self.expect(')'); return '(%s)' % c, t
        if (k, v) == ('op', '-'):
            c, t = self.p_primary(); return '(-%s)' % c, t
        if k == 'id':
            u = self.u
            if self.peek() == ('op', '('):
                if v in u.dims:
                    args = self.p_args()
                    return u_index(u, v, [a[0] for a in args]), u.typeof(v)
                if v in INTRINSICS:
                    args = self.p_args(); f, t = INTRINSICS[v]
                    return '%s(%s)' % (f, args[0][0]), t
                if v == 'MOD':
                    args = self.p_args()
                    return '((%s) %% (%s))' % (args[0][0], args[1][0]), 'i'
                if v in EXTERNAL_FUNCS:
                    args = self.p_args()
                    return '%s_(%s)' % (v.lower(), ', '.join(call_arg(u, a) for a in args)), EXTERNAL_FUNCS[v]
                raise SyntaxError('unknown function/array %s in unit %s' % (v, u.name))
            return u_scalar(u, v), u.typeof(v)
        raise SyntaxError('unexpected token %r' % ((k, v),))


def cname(u, name):
    if name in u.common:
        return u.common[name]
    return name


def u_scalar(u, name):
    if name in u.params:
        return name
    if name in u.dims:
        return cname(u, name)           # whole array (pointer)
    if name in u.args:
        if u.typeof(name) == 'c':
            return '%s[0]' % name
        return '(*%s)' % name
    return cname(u, name)


def u_index(u, name, idx):
    dims = u.dims[name]
    if len(idx) == 1:
        return '%s[(%s)-1]' % (cname(u, name), idx[0])
    if len(idx) == 2:
        ld = dim_c(u, dims[0])
        return '%s[((%s)-1) + (size_t)((%s)-1)*(%s)]' % (cname(u, name), idx[0], idx[1], ld)
    if len(idx) == 3:
        d0 = dim_c(u, dims[0]); d1 = dim_c(u, dims[1])
        return '%s[((%s)-1) + (size_t)((%s)-1)*(%s) + (size_t)((%s)-1)*(%s)*(%s)]' % (
            cname(u, name), idx[0], idx[1], d0, idx[2], d0, d1)
    raise SyntaxError('rank>3 arrays unsupported')


def dim_c(u, dimtoks):
    if dimtoks == [('op', '*')]:
        return '1'
    return ExprParser(u, dimtoks).parse()[0]


def call_arg(u, a):
    c, t, toks = a
    if len(toks) == 1 and toks[0][0] == 'id':
        n = toks[0][1]
        if n in u.params:
            return '(%s[]){%s}' % ('double' if u.typeof(n) == 'd' else 'int', n)
        if n in u.dims:
            return cname(u, n)
        if n in u.args:
            return n
        return '&%s' % cname(u, n)
    if len(toks) >= 4 and toks[0][0] == 'id' and toks[0][1] in u.dims and toks[1] == ('op', '(') and toks[-1] == ('op', ')'):
        # array element (only if the outer parens match as one group)
        depth = 0; ok = True
        for i, tk in enumerate(toks[1:], 1):
            if tk == ('op', '('):
                depth += 1
            elif tk == ('op', ')'):
                depth -= 1
                if depth == 0 and i != len(toks) - 1:
                    ok = False; break
        if ok:
            return '&%s' % c
    if t == 'c':
        return '"%s"' % c.strip("'\"")
    return '(%s[]){%s}' % ('double' if t == 'd' else 'int', c)


def split_top(toks, sep=','):
    parts = []; cur = []; depth = 0
    for tk in toks:
        if tk == ('op', '('):
            depth += 1
        elif tk == ('op', ')'):
            depth -= 1
        if tk == ('op', sep) and depth == 0:
            parts.append(cur); cur = []
        else:
            cur.append(tk)
    parts.append(cur)
    return parts


def parse_decl_list(u, toks, typ=None):
    for part in split_top(toks):
        if not part:
            continue
        name = part[0][1]
        # CHARACTER*1 X  -> tokens: '*','1', name ...
        if part[0] == ('op', '*'):
            part = part[2:]; name = part[0][1]
        if part[0] == ('op', '('):      # CHARACTER(*) PLIK
            close = part.index(('op', ')')); part = part[close + 1:]; name = part[0][1]
        if typ:
            u.types[name] = typ
        if len(part) > 1 and part[1] == ('op', '('):
            u.dims[name] = split_top(part[2:-1])


def parse_units(text):
    units = []; u = None
    for label, st in logical_lines(text):
        up = st.upper()
        if up.startswith('SUBROUTINE'):
            m = re.match(r'SUBROUTINE\s+(\w+)\s*\((.*)\)', up.replace(' ', '').replace('SUBROUTINE', 'SUBROUTINE ', 1))
            name = m.group(1); args = [a for a in m.group(2).split(',') if a]
            u = Unit(name, args); units.append(u); continue
        if up.startswith('BLOCK DATA') or up.startswith('BLOCKDATA'):
            u = Unit('__BLOCKDATA__', []); units.append(u); continue
        if u is None:
            continue
        u.body.append((label, st))
    return units


def emit_unit(u, out, commons):
    body = []
    decl_done = False
    # ---- pass 1: declarations
    stmts = []
    for label, st in u.body:
        up = st.upper().strip()
        toks = None
        if up.startswith('IMPLICIT'):
            continue
        if up.startswith('DIMENSION'):
            parse_decl_list(u, tokenize(st[9:])); continue
        if up.startswith('DOUBLE PRECISION'):
            parse_decl_list(u, tokenize(st[16:]), 'd'); continue
        if up.startswith('INTEGER'):
            parse_decl_list(u, tokenize(st[7:]), 'i'); continue
        if up.startswith('LOGICAL'):
            parse_decl_list(u, tokenize(st[7:]), 'l'); continue
        if up.startswith('CHARACTER'):
            parse_decl_list(u, tokenize(st[9:]), 'c'); continue
        if up.startswith('EXTERNAL'):
            continue
        if up.startswith('PARAMETER'):
            inner = tokenize(st[st.index('(') + 1: st.rindex(')')])
            for part in split_top(inner):
                u.params[part[0][1]] = part[2:]
            continue
        if up.startswith('COMMON'):
            m = re.match(r'COMMON\s*/\s*(\w+)\s*/(.*)', st, re.I)
            blk = m.group(1).upper()
            toks = tokenize(m.group(2))
            before = set(u.dims)
            parse_decl_list(u, toks)
            for part in split_top(toks):
                n = part[0][1]
                g = 'common_%s_%s' % (blk, n)
                u.common[n] = g
                commons[g] = (u.typeof(n), [dim_c(u, d) for d in u.dims.get(n, [])])
            continue
        if up.startswith('DATA'):
            m = re.match(r'DATA\s+(\w+)\s*/', st, re.I)
            u.data_zero.add(m.group(1).upper()); continue
        stmts.append((label, st))
    if u.name == '__BLOCKDATA__':
        return
    if u.name in SKIP_UNITS:
        return
    ctype = {'d': 'double', 'i': 'int', 'l': 'int', 'c': 'const char'}
    sig = ', '.join('%s *%s' % (ctype[u.typeof(a)], a) for a in u.args)
    out.append('void %s_(%s)\n{' % (u.name.lower(), sig))
    for n, toks in u.params.items():
        c, t = ExprParser(u, toks).parse()
        out.append('  const %s %s = %s;' % ('double' if u.typeof(n) == 'd' else 'int', n, c))
    # collect identifiers used
    used = set()
    for label, st in stmts:
        try:
            for k, v in tokenize(re.sub(r"^\s*\d+\s+", '', st)):
                if k == 'id':
                    used.add(v)
        except SyntaxError:
            pass
    keywords = {'DO', 'IF', 'THEN', 'ELSE', 'ENDIF', 'ENDDO', 'CALL', 'GOTO', 'CONTINUE', 'RETURN', 'END',
                'WRITE', 'OPEN', 'CLOSE', 'ELSEIF', 'GO', 'TO', 'FILE', 'ACCESS', 'FORM'}
    frees = []
    for n in sorted(u.dims):
        if n in u.args or n in u.common:
            continue
        dims = [dim_c(u, d) for d in u.dims[n]]
        size = '*'.join('(size_t)(%s)' % d for d in dims)
        const = all(re.match(r'^[\d()\s*]+$', d) for d in dims)
        if const:
            out.append('  static %s %s[%s];' % (ctype[u.typeof(n)], n, '*'.join(d.strip('()') for d in dims)))
        else:
            out.append('  %s *%s = (%s*)calloc(%s, sizeof(%s));' % (ctype[u.typeof(n)], n, ctype[u.typeof(n)], size, ctype[u.typeof(n)]))
            frees.append(n)
    for n in sorted(used):
        if n in u.args or n in u.dims or n in u.params or n in u.common or n in keywords:
            continue
        if n in INTRINSICS or n in EXTERNAL_FUNCS or n in SKIP_CALLS or n == 'MOD':
            continue
        if n in CALLED or n in EXTERNAL_SUBS:
            continue
        out.append('  %s %s = 0;' % (ctype[u.typeof(n)], n))
    # ---- pass 2: executable statements
    do_stack = []      # labels ('' for block DO)
    ind = '  '

    def expr(s):
        return ExprParser(u, tokenize(s)).parse()[0]

    def simple(st):
        """translate a simple (non-block) statement to C"""
        up = st.upper().strip()
        if up.startswith('GOTO') or up.startswith('GO TO'):
            return 'goto L%s;' % re.sub(r'\D', '', up)
        if up == 'RETURN':
            return 'goto L_return;'
        if up == 'CONTINUE':
            return ';'
        if up.startswith('WRITE') or up.startswith('OPEN') or up.startswith('CLOSE') or up.startswith('PRINT'):
            return ';'
        if up.startswith('CALL'):
            toks = tokenize(st.strip()[4:])
            name = toks[0][1]
            if name in SKIP_CALLS:
                return ';'
            p = ExprParser(u, toks[1:])
            args = p.p_args() if len(toks) > 1 else []
            return '%s_(%s);' % (name.lower(), ', '.join(call_arg(u, a) for a in args))
        # assignment
        toks = tokenize(st)
        depth = 0; eq = None
        for i, tk in enumerate(toks):
            if tk == ('op', '('):
                depth += 1
            elif tk == ('op', ')'):
                depth -= 1
            elif tk == ('op', '=') and depth == 0:
                eq = i; break
        if eq is None:
            raise SyntaxError('cannot translate: %r' % st)
        lhs = ExprParser(u, toks[:eq]).parse()[0]
        rhs = ExprParser(u, toks[eq + 1:]).parse()[0]
        return '%s = %s;' % (lhs, rhs)

    for label, st in stmts:
        up = st.upper().strip()
        pre = ''
        if label:
            pre = 'L%s: ;' % label
        if up == 'END':
            break
        m = re.match(r'DO\s*(\d+)?\s*,?\s*(\w+)\s*=\s*(.*)$', st.strip(), re.I)
        if m and re.match(r'DO[\s\d]', up):
            lab, var, rest = m.group(1), m.group(2).upper(), m.group(3)
            parts = split_top(tokenize(rest))
            a = ExprParser(u, parts[0]).parse()[0]; b = ExprParser(u, parts[1]).parse()[0]
            s = ExprParser(u, parts[2]).parse()[0] if len(parts) > 2 else '1'
            v = u_scalar(u, var)
            if pre:
                out.append(ind + pre)
            if s == '1':
                out.append(ind + 'for (%s = %s; %s <= %s; %s++) {' % (v, a, v, b, v))
            else:
                out.append(ind + 'for (%s = %s; (%s) > 0 ? %s <= %s : %s >= %s; %s += %s) {' % (v, a, s, v, b, v, b, v, s))
            do_stack.append(lab or '')
            ind += '  '
            continue
        if up in ('ENDDO', 'END DO'):
            if pre:
                out.append(ind + pre)
            assert do_stack and do_stack[-1] == '', 'ENDDO mismatch in %s' % u.name
            do_stack.pop(); ind = ind[:-2]; out.append(ind + '}')
            continue
        m = re.match(r'IF\s*\((.*)\)\s*THEN$', st.strip(), re.I)
        if m:
            if pre:
                out.append(ind + pre)
            out.append(ind + 'if (%s) {' % expr(m.group(1))); ind += '  '; continue
        m = re.match(r'ELSE\s*IF\s*\((.*)\)\s*THEN$', st.strip(), re.I)
        if m:
            ind = ind[:-2]; out.append(ind + '} else if (%s) {' % expr(m.group(1))); ind += '  '; continue
        if up == 'ELSE':
            ind = ind[:-2]; out.append(ind + '} else {'); ind += '  '; continue
        if up in ('ENDIF', 'END IF'):
            ind = ind[:-2]; out.append(ind + '}'); continue
        if up.startswith('IF'):
            # logical IF: find matching paren
            s = st.strip(); i = s.index('('); depth = 0
            for j in range(i, len(s)):
                if s[j] == '(':
                    depth += 1
                elif s[j] == ')':
                    depth -= 1
                    if depth == 0:
                        break
            cond = s[i + 1:j]; rest = s[j + 1:].strip()
            out.append(ind + pre + ' if (%s) { %s }' % (expr(cond), simple(rest)))
        else:
            out.append(ind + pre + ' ' + simple(st))
        # close labelled DO loops terminated by this label
        while label and do_stack and do_stack[-1] == label:
            do_stack.pop(); ind = ind[:-2]; out.append(ind + '}')
    out.append('  L_return: ;')
    for n in frees:
        out.append('  free(%s);' % n)
    out.append('}\n')


CALLED = set()


def transpile(text, only=None):
    units = parse_units(text)
    if only is not None:
        units = [u for u in units if u.name in only]
    for u in units:
        CALLED.add(u.name)
    commons = {}
    bodies = []
    for u in units:
        emit_unit(u, bodies, commons)
    head = ['/* GENERATED by oracle/f77c.py from the reference Fortran -- do not commit, do not edit */',
            '#include <math.h>', '#include <stdlib.h>', '#include <stddef.h>',
            'double ddot_(int*, double*, int*, double*, int*);',
            'int ilaenv_(int*, const char*, const char*, int*, int*, int*, int*);',
            'void dgemm_(const char*, const char*, int*, int*, int*, double*, double*, int*, double*, int*, double*, double*, int*);',
            'void dgetrf_(int*, int*, double*, int*, int*, int*);',
            'void dgetri_(int*, double*, int*, int*, double*, int*, int*);']
    for g, (t, dims) in commons.items():
        head.append('static %s %s[%s];' % ('double' if t == 'd' else 'int', g, '*'.join(d.strip('()') for d in dims) or '1'))
    # prototypes
    protos = []
    for line in bodies:
        if line.startswith('void ') and line.endswith('{'):
            protos.append(line[:-2] + ';')
    return '\n'.join(head + protos + [''] + bodies) + '\n'


if __name__ == '__main__':
    src, dst = sys.argv[1], sys.argv[2]
    with open(src, encoding='utf-8', errors='replace') as fh:
        text = fh.read()
    with open(dst, 'w') as fh:
        fh.write(transpile(text))
