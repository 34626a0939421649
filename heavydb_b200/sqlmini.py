"""Tiny SQL front-end for tests: turns the reference's own test query strings (Tests/ExecuteTest.cpp) into the
RelAlgExecutionUnit mirror, for the subset of the path:

    SELECT <col | COUNT(*) | COUNT(c) | SUM(c) | MIN(c) | MAX(c) | AVG(c)>, ...
    FROM <table> [[LEFT] JOIN <inner> ON <table>.<c> = <inner>.<c>] [WHERE <c OP literal | c IS [NOT] NULL | c [NOT] IN (l, ...) | c BETWEEN l AND l | NOT <factor>>
                  {AND|OR} ... with parentheses] [GROUP BY c {, c}]
    [ORDER BY <position | target text | target alias> [ASC|DESC] [NULLS FIRST|LAST] {, ...}] [LIMIT n] [OFFSET m]

It plays the role Calcite + RelAlgTranslator play in the reference (kept, out of scope) and is test infrastructure.
Like RelAlgTranslator/QualsConjunctiveForm, a top-level AND is split into separate quals, and a `col OP const`
conjunct is a "simple qual" (Analyzer::BinOper::normalize_simple_predicate, QueryEngine/RelAlgExecutor.cpp
translation of filters: simple_quals vs quals).
"""
from __future__ import annotations

import re
from typing import List, Tuple

from heavydb_b200 import abi

_TOK = re.compile(r"\s*('(?:[^']|'')*'|<>|<=|>=|!=|[(),*<>=]|[A-Za-z_][A-Za-z_0-9]*(?:\.[A-Za-z_][A-Za-z_0-9]*)?|-?\d+\.\d*(?:[eE][-+]?\d+)?|-?\d+)")
_OPS = {"=": abi.kEQ, "<>": abi.kNE, "!=": abi.kNE, "<": abi.kLT, ">": abi.kGT, "<=": abi.kLE, ">=": abi.kGE}
_AGGS = {"COUNT": abi.kCOUNT, "SUM": abi.kSUM, "MIN": abi.kMIN, "MAX": abi.kMAX, "AVG": abi.kAVG}


def _tokens(s: str) -> List[str]:
    s = s.strip().rstrip(";")
    out, pos = [], 0
    while pos < len(s):
        m = _TOK.match(s, pos)
        if not m:
            raise ValueError(f"cannot tokenize at: {s[pos:]!r}")
        out.append(m.group(1))
        pos = m.end()
    return out


class _P:
    def __init__(self, toks, table: abi.Table, names: List[str], bigint_count: bool, inner=None, dicts=None):
        self.t, self.i = toks, 0
        self.dicts = {k.lower(): v for k, v in (dicts or {}).items()}   # column name -> strings in dictionary-id order
        self.b = abi.UnitBuilder(table)
        self.names = [n.lower() for n in names]
        self.bigint_count = bigint_count
        # one INNER join level: inner = (abi.Table, [column names]); set before any column is resolved
        self.inner_names = [n.lower() for n in inner[1]] if inner else []
        self.outer_alias = self.inner_alias = None
        if inner:
            self.b.inner = inner[0]

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def eat(self, expect=None):
        tok = self.peek()
        if tok is None or (expect is not None and tok.upper() != expect):
            raise ValueError(f"expected {expect}, got {tok}")
        self.i += 1
        return tok

    def colref(self, name):
        """(column id, rte_idx).  `alias.col` is resolved by alias, a bare name in the outer table first."""
        name = name.lower()
        if "." in name:
            alias, col = name.split(".", 1)
            if self.inner_alias is not None and alias == self.inner_alias:
                return self.inner_names.index(col), 1
            return self.names.index(col), 0
        if name in self.names:
            return self.names.index(name), 0
        return self.inner_names.index(name), 1

    def colid(self, name):
        c, rte = self.colref(name)
        assert rte == 0, f"{name}: inner-table column where only outer columns are supported"
        return c

    # cond := term {OR term}; term := factor {AND factor}; factor := '(' cond ')' | col OP literal
    def cond(self):
        e = self.term()
        while self.peek() and self.peek().upper() == "OR":
            self.eat()
            e = self.b.binop(abi.kOR, e, self.term())
        return e

    def term(self):
        e = self.factor()
        while self.peek() and self.peek().upper() == "AND":
            self.eat()
            e = self.b.binop(abi.kAND, e, self.factor())
        return e

    def literal_cmp(self, col, op, lit, rte=0):
        tbl = self.b.inner if rte else self.b.table
        if lit.startswith("'"):
            # a string literal against a dictionary-encoded column: the reference translates it to the column's dictionary id
            # (StringDictionaryProxy::getIdOfString; an unknown string is INVALID_STR_ID = -1, which no row holds) and compares ids
            # — for = and <> only; an ordering comparison needs the dictionary itself (CodeGenerator::codegenCmp -> string ops)
            name = (self.inner_names if rte else self.names)[col]
            if tbl.col_types[col][0] not in (abi.kTEXT, abi.kVARCHAR, abi.kCHAR) or name not in self.dicts:
                raise ValueError(f"string literal {lit} against a column without a dictionary")
            if op not in (abi.kEQ, abi.kNE):
                raise ValueError("string ordering comparisons need the dictionary: outside the path")
            text = lit[1:-1].replace("''", "'")
            sid = self.dicts[name].index(text) if text in self.dicts[name] else -1
            return self.b.cmp(col, op, sid, abi.kBIGINT, rte)
        if tbl.col_types[col][0] in abi.DECIMAL_TYPES:
            # the analyzer folds the literal to the common DECIMAL type: same scale as the column when the literal has no
            # more fraction digits than the column (Constant::do_cast); otherwise the COLUMN would be cast — not this path
            import decimal
            scale = tbl.col_scales.get(col, 0)
            v = decimal.Decimal(lit).scaleb(scale)
            if v != v.to_integral_value():
                raise ValueError(f"literal {lit} has more fraction digits than DECIMAL scale {scale}")
            return self.b.cmp(col, op, int(v), tbl.col_types[col][0], rte, scale=scale)
        if tbl.col_types[col][0] == abi.kFLOAT:
            # common_numeric_type(FLOAT, <int / decimal literal>) = FLOAT: the literal is folded to a FLOAT Datum
            return self.b.cmp(col, op, float(lit), abi.kFLOAT, rte)
        if re.fullmatch(r"-?\d+", lit):
            return self.b.cmp(col, op, int(lit), abi.kBIGINT, rte)
        return self.b.cmp(col, op, float(lit), abi.kDOUBLE, rte)

    def factor(self):
        if self.peek() == "(":
            self.eat()
            e = self.cond()
            self.eat(")")
            return e
        if self.peek().upper() == "NOT":          # Analyzer::UOper(kNOT, ...)
            self.eat()
            return self.b.uoper(abi.kNOT, self.factor())
        col, rte = self.colref(self.eat())
        nxt = self.peek().upper()
        if nxt == "IS":                            # c IS NULL -> UOper(kISNULL, c); IS NOT NULL -> NOT(ISNULL), like RelAlgTranslator
            self.eat()
            neg = self.peek().upper() == "NOT"
            if neg:
                self.eat()
            self.eat("NULL")
            e = self.b.uoper(abi.kISNULL, self.b.col(col, rte))
            return self.b.uoper(abi.kNOT, e) if neg else e
        neg = False
        if nxt == "NOT":
            self.eat()
            neg = True
            nxt = self.peek().upper()
        if nxt == "IN":                            # c IN (a, b, ...) -> OR of equalities (InValues codegen for short lists)
            self.eat()
            self.eat("(")
            e = self.literal_cmp(col, abi.kEQ, self.eat(), rte)
            while self.peek() == ",":
                self.eat()
                e = self.b.binop(abi.kOR, e, self.literal_cmp(col, abi.kEQ, self.eat(), rte))
            self.eat(")")
            return self.b.uoper(abi.kNOT, e) if neg else e
        if nxt == "BETWEEN":                       # Calcite expands BETWEEN to >= AND <=
            self.eat()
            lo = self.eat()
            self.eat("AND")
            hi = self.eat()
            e = self.b.binop(abi.kAND, self.literal_cmp(col, abi.kGE, lo, rte), self.literal_cmp(col, abi.kLE, hi, rte))
            return self.b.uoper(abi.kNOT, e) if neg else e
        op = _OPS[self.eat()]
        rhs = self.eat()
        if re.fullmatch(r"[A-Za-z_][A-Za-z_0-9.]*", rhs):     # column OP column
            c2, rte2 = self.colref(rhs)
            if self.dicts:   # ids of two dictionaries are not comparable: what the bridge's on_path() checks with getStringDictKey()
                n1 = (self.inner_names if rte else self.names)[col]
                n2 = (self.inner_names if rte2 else self.names)[c2]
                if n1 in self.dicts and n2 in self.dicts and self.dicts[n1] is not self.dicts[n2]:
                    raise ValueError(f"{n1} and {n2} use different dictionaries: string comparison, outside the path")
            return self.b.binop(op, self.b.col(col, rte), self.b.col(c2, rte2))
        return self.literal_cmp(col, op, rhs, rte)

    def target_text(self):
        """Canonical text of the target expression starting at the cursor (does not build nodes)."""
        j = self.i
        tok = self.t[j]
        if tok.upper() in _AGGS and j + 1 < len(self.t) and self.t[j + 1] == "(":
            k = self.t.index(")", j)
            return "".join(self.t[j:k + 1]).upper(), k + 1
        return tok.upper(), j + 1

    def target(self):
        tok = self.eat()
        if tok.upper() in _AGGS and self.peek() == "(":
            self.eat("(")
            if self.peek() == "*":
                self.eat()
                self.eat(")")
                return self.b.agg(abi.kCOUNT, None, self.bigint_count)
            distinct = False
            if self.peek().upper() == "DISTINCT":
                self.eat()
                distinct = True
            col, rte = self.colref(self.eat())
            self.eat(")")
            return self.b.agg(_AGGS[tok.upper()], col, self.bigint_count, rte, is_distinct=distinct)
        return self.b.col(*self.colref(tok))


def _conjuncts(b: abi.UnitBuilder, e: int) -> List[int]:
    n = b.nodes[e]
    if n.kind == abi.EXPR_BIN_OPER and n.op == abi.kAND:
        return _conjuncts(b, n.left) + _conjuncts(b, n.right)
    return [e]


def parse(sql: str, table: abi.Table, names: List[str], bigint_count: bool = False, inner=None, dicts=None) -> abi.BuiltUnit:
    """inner = (abi.Table, [names]) of the table named after JOIN (one concatenated fragment); dicts = {column: [strings in id
    order]} for string literals against dictionary-encoded columns."""
    toks = _tokens(sql)
    p = _P(toks, table, names, bigint_count, inner, dicts)
    ups = [t.upper() for t in toks]
    fi = ups.index("FROM")
    p.outer_alias = toks[fi + 1].lower()
    ji = fi + 2
    if ji < len(toks) and ups[ji] == "LEFT":
        p.b.join_type = 1          # known before any column of the inner table is resolved (they become nullable)
        ji += 1
    if ji < len(toks) and ups[ji] == "JOIN":
        assert inner is not None, "JOIN needs the inner table"
        p.inner_alias = toks[ji + 1].lower()
    p.eat("SELECT")
    texts, targets, aliases = [], [], {}

    def one_target():
        texts.append(p.target_text()[0])
        targets.append(p.target())
        if p.peek() and p.peek().upper() == "AS":     # <target> AS alias — usable in ORDER BY
            p.eat()
            aliases[p.eat().upper()] = len(targets)
        elif p.peek() and p.peek() != "," and p.peek().upper() != "FROM" and re.fullmatch(r"[A-Za-z_][A-Za-z_0-9]*", p.peek()):
            aliases[p.eat().upper()] = len(targets)

    one_target()
    while p.peek() == ",":
        p.eat()
        one_target()
    p.eat("FROM")
    p.eat()  # table name
    if p.peek() and p.peek().upper() == "LEFT":
        p.eat()
    if p.peek() and p.peek().upper() == "JOIN":   # join_quals[0] = {a = b}, INNER or LEFT
        p.eat()
        p.eat()  # inner table name
        p.eat("ON")
        (c1, r1), _, (c2, r2) = p.colref(p.eat()), p.eat("="), p.colref(p.eat())
        assert {r1, r2} == {0, 1}, "ON must compare an outer with an inner column"
        p.b.join(inner[0], c1 if r1 == 0 else c2, c2 if r1 == 0 else c1, p.b.join_type)
    if p.peek() and p.peek().upper() == "WHERE":
        p.eat()
        e = p.cond()
        for c in _conjuncts(p.b, e):
            n = p.b.nodes[c]
            simple = n.kind == abi.EXPR_BIN_OPER and n.op not in (abi.kAND, abi.kOR)
            if simple and p.b.nodes[n.right].kind != abi.EXPR_CONSTANT:
                simple = False          # column OP column is never a simple qual
            if simple:
                # an integer column against an fp literal is `CAST(col AS DOUBLE) OP lit` in the reference; only
                # int->int / timestamp casts survive BinOper::normalize_simple_predicate, so it is NOT a simple qual
                col_fp = p.b.nodes[n.left].type in (abi.kDOUBLE, abi.kFLOAT)
                lit_fp = p.b.nodes[n.right].type in (abi.kDOUBLE, abi.kFLOAT)
                simple = col_fp == lit_fp
            p.b.add_qual(c, simple=simple)
    if p.peek() and p.peek().upper() == "GROUP":
        p.eat()
        p.eat("BY")
        p.b.group_by(*p.colref(p.eat()))
        while p.peek() == ",":
            p.eat()
            p.b.group_by(*p.colref(p.eat()))
    if p.peek() and p.peek().upper() == "ORDER":
        p.eat()
        p.eat("BY")
        while True:
            if re.fullmatch(r"\d+", p.peek()):
                tle = int(p.eat())
            elif p.peek().upper() in aliases:
                tle = aliases[p.eat().upper()]
            else:
                text, nxt = p.target_text()
                p.i = nxt
                tle = texts.index(text) + 1
            desc, nulls_first = False, None
            if p.peek() and p.peek().upper() in ("ASC", "DESC"):
                desc = p.eat().upper() == "DESC"
            if p.peek() and p.peek().upper() == "NULLS":
                p.eat()
                nulls_first = p.eat().upper() == "FIRST"
            p.b.order_by(tle, desc, nulls_first)
            if p.peek() != ",":
                break
            p.eat()
    if p.peek() and p.peek().upper() == "LIMIT":
        p.eat()
        p.b.limit = int(p.eat())
    if p.peek() and p.peek().upper() == "OFFSET":
        p.eat()
        p.b.offset = int(p.eat())
    if p.peek() is not None:
        raise ValueError(f"trailing tokens: {p.t[p.i:]}")
    for t in targets:
        p.b.target(t)
    return p.b.build()
