"""The C-ABI library loads on a CPU-only box, exports every symbol include/b2q.h declares, and the ctypes mirror
(heavydb_b200/abi.py) has the C compiler's struct sizes.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from heavydb_b200 import abi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b2q.h")


@pytest.fixture(scope="module")
def libpath():
    return build.build()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2q_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(libpath):
    lib = C.CDLL(libpath)
    names = declared_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in b2q.h but not exported: {missing}"


def test_struct_sizes_match_c_compiler(tmp_path):
    structs = {
        "B2QTypeInfo": abi.TypeInfo, "B2QExpr": abi.Expr, "B2QExecUnit": abi.ExecUnit, "B2QChunkStats": abi.ChunkStats,
        "B2QFragmentInfo": abi.FragmentInfo, "B2QTableInfo": abi.TableInfo, "B2QCompilationOptions": abi.CompilationOptions,
        "B2QExecutionOptions": abi.ExecutionOptions, "B2QTargetInfo": abi.TargetInfo, "B2QPlan": abi.Plan,
        "B2QTargetValue": abi.TargetValue, "B2QParams": abi.Params,
    }
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "b2q.h"\nint main(){\n'
    for n in structs:
        prog += f'printf("{n} %zu\\n", sizeof({n}));\n'
    prog += 'printf("off_init_vals %zu\\n", offsetof(B2QPlan, init_vals));\n'
    prog += 'printf("off_targets %zu\\n", offsetof(B2QPlan, targets));\n'
    prog += 'printf("off_dval %zu\\n", offsetof(B2QExpr, dval));\nreturn 0;}\n'
    src = tmp_path / "sz.c"
    src.write_text(prog)
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = dict(line.split() for line in subprocess.check_output([str(exe)]).decode().splitlines())
    for n, cls in structs.items():
        assert int(out[n]) == C.sizeof(cls), n
    assert int(out["off_init_vals"]) == abi.Plan.init_vals.offset
    assert int(out["off_targets"]) == abi.Plan.targets.offset
    assert int(out["off_dval"]) == abi.Expr.dval.offset


def test_no_device_is_an_error_not_a_fallback(libpath):
    """On a box without a GPU every compute entry must fail loudly (B2Q_ERR_NO_DEVICE); with a GPU this test is moot."""
    from heavydb_b200 import executor
    import ref_tables as rt
    import sqlmini
    if executor.lib().b2q_device_count() > 0:
        pytest.skip("a CUDA device is present")
    table = rt.make_table(rt.test_rows())
    unit = sqlmini.parse("SELECT COUNT(*) FROM test;", table, rt.TEST_NAMES)
    with pytest.raises(executor.NoDeviceError):
        executor.Executor().executeWorkUnit(0, True, table, unit)


def test_malformed_units_are_refused_not_followed():
    """Operands must be earlier nodes of the expression array: a self-referencing / forward-referencing node (a cycle
    would recurse forever) and negative fragment row counts come back as INVALID_ARGUMENT from the host-only planner."""
    import ref_tables as rt
    import sqlmini
    from heavydb_b200 import executor
    table = rt.make_table(rt.test_rows())
    unit = sqlmini.parse("SELECT x, COUNT(*) FROM test WHERE y > 42 OR z < 100 GROUP BY x;", table, rt.TEST_NAMES)
    u = unit.unit
    ors = [i for i in range(u.num_exprs) if u.exprs[i].kind == abi.EXPR_BIN_OPER and u.exprs[i].op == abi.kOR]
    assert ors
    saved = u.exprs[ors[0]].right
    for bad in (ors[0], u.num_exprs, -7):
        u.exprs[ors[0]].right = bad
        with pytest.raises(executor.QueryExecutionError) as ei:
            executor.Executor().plan(unit, table)
        assert ei.value.code == abi.ERR_INVALID_ARGUMENT
    u.exprs[ors[0]].right = saved
    executor.Executor().plan(unit, table)
    table.fragments[0].num_tuples = -1
    with pytest.raises(executor.QueryExecutionError) as ei:
        executor.Executor().plan(unit, table)
    assert ei.value.code == abi.ERR_INVALID_ARGUMENT
