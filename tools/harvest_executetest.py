#!/usr/bin/env python
"""TEST INFRASTRUCTURE (fixture generator; never imported by the product, tests, smoke() or bench.py — it is the committed script that
made tests/golden/executetest_harvest.json).  Harvest the reference's own golden query strings for this path (run HERE, where /root/reference exists; the result is
committed as tests/golden/executetest_harvest.json and replayed by tests/test_oracle_golden_harvest.py without the reference).

    python tools/harvest_executetest.py > tests/golden/executetest_harvest.json

Tests/ExecuteTest.cpp holds ~1350 `c("SELECT ...", dt)` comparisons against SQLite.  A string is kept when
  * it reads only table `test` and only the columns tests/ref_full_table.py models (the numeric, dictionary-string and FIXED columns
    of the golden table),
  * heavydb_b200.sqlmini parses it (one table, AND/OR/NOT of column-vs-constant / column-vs-column / IS NULL / IN / BETWEEN,
    GROUP BY columns, COUNT / SUM / MIN / MAX / AVG / COUNT(DISTINCT) of a column, ORDER BY / LIMIT / OFFSET),
  * the oracle plans it (the path's own refusals drop the rest), and
  * SQLite evaluates the same string.
Only the query string and the reference line it came from are stored: the expected rows are recomputed with SQLite at test time,
exactly as the reference's SQLiteComparator does (ExecuteTest.cpp:383-520)."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib  # noqa: E402
import ref_full_table as ft  # noqa: E402
import ref_tables as rt  # noqa: E402
import sqlmini  # noqa: E402
from heavydb_b200 import abi  # noqa: E402

SRC = "/root/reference/Tests/ExecuteTest.cpp"


def main():
    text = open(SRC).read()
    rows = ft.full_rows()
    table = ft.make_table(rows)
    con = ft.make_sqlite(rows)
    seen, out, stats = set(), [], {"strings": 0, "table_test": 0, "parsed": 0, "planned": 0, "kept": 0}
    # c("..." "..." , dt) with adjacent literals concatenated, or c(R"(...)", dt); comparisons wrapped in EXPECT_THROW are negative tests
    pat = re.compile(r'\bc\(\s*(?:R"\((.*?)\)"|((?:"(?:[^"\\]|\\.)*"\s*)+))\s*,', re.S)
    for _ in (0,):
        for m in pat.finditer(text):
            if re.search(r"EXPECT_(ANY_)?THROW\(\s*$", text[max(0, m.start() - 40):m.start()]):
                continue
            q = m.group(1) if m.group(1) is not None else "".join(re.findall(r'"((?:[^"\\]|\\.)*)"', m.group(2)))
            q = re.sub(r"\s+", " ", q).strip()
            if not q.upper().startswith("SELECT"):
                continue
            ln = text.count("\n", 0, m.start()) + 1
            stats["strings"] += 1
            if not re.search(r"\bFROM test\b", q) or re.search(r"\b(JOIN|UNION|OVER|CASE|HAVING|EXTRACT|CAST|LIKE|DISTINCT ON)\b|,\s*test\b|\(SELECT", q, re.I):
                continue
            stats["table_test"] += 1
            sql = q if q.endswith(";") else q + ";"
            if sql in seen:
                continue
            try:
                unit = sqlmini.parse(sql, table, ft.FULL_NAMES, dicts=ft.DICTS)
            except Exception:
                continue
            stats["parsed"] += 1
            try:
                res = oracle_lib.execute(unit, table, entry_guess=48, has_card=True, num_threads=2)
            except oracle_lib.OracleError:
                continue
            stats["planned"] += 1
            try:
                con.execute(sql.rstrip(";")).fetchall()
            except Exception:
                continue
            del res
            seen.add(sql)
            out.append({"sql": sql, "line": ln})
    stats["kept"] = len(out)
    json.dump({"source": "Tests/ExecuteTest.cpp", "how": "tools/harvest_executetest.py", "stats": stats, "queries": out}, sys.stdout, indent=0)
    print()
    print(stats, file=sys.stderr)


if __name__ == "__main__":
    main()
