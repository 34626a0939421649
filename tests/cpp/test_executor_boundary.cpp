/*
 * Boundary-level test in the shape of the reference's Tests/GroupByTest.cpp:73-152 (`PerfectHashNoFallback`) and
 * :173-338 (`BaselineFallbackTest`, `BaselineNoFilters`): build a RelAlgExecutionUnit by hand, call
 * executor->executeWorkUnit(...) directly, check rowCount()/getNextRow().  Like the reference the first test groups
 * by a dictionary-encoded string column (the chunk holds the ids).
 *
 * Table: (x, big, sparse, str) = (1, 10, 9e12, 'hi'), (2, 20, 7, 'bye')   — GroupByTest.cpp:60-63 inserts (1,'hi'), (2,'bye').
 */
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "b2q_executor.hpp"

using namespace b2q;

#define CHECK(c)                                                        \
  do {                                                                  \
    if (!(c)) { std::fprintf(stderr, "CHECK failed: %s (line %d)\n", #c, __LINE__); std::exit(1); } \
  } while (0)

int main() {
  std::vector<int32_t> x{1, 2};
  std::vector<int64_t> big{10, 20};
  std::vector<int64_t> sparse{9000000000000LL, 7};
  InputTableInfo info;
  std::vector<int32_t> str_ids{0, 1}; /* 'hi' -> 0, 'bye' -> 1 */
  info.col_types = {SQLTypeInfo(kINT, true), SQLTypeInfo(kBIGINT, false), SQLTypeInfo(kBIGINT, true), SQLTypeInfo(kTEXT, false)};
  info.memory_level = MemoryLevel::CPU_LEVEL; /* host chunks: fetched to the GPU inside the call */
  FragmentInfo f;
  f.numTuples = 2;
  f.col_buffers = {x.data(), big.data(), sparse.data(), str_ids.data()};
  f.chunkStats.resize(4);
  f.chunkStats[3].int_min = 0; f.chunkStats[3].int_max = 1;
  f.chunkStats[0].int_min = 1; f.chunkStats[0].int_max = 2;
  f.chunkStats[1].int_min = 10; f.chunkStats[1].int_max = 20;
  f.chunkStats[2].int_min = 7; f.chunkStats[2].int_max = 9000000000000LL;
  info.fragments.push_back(f);

  auto executor = std::make_shared<Executor>();
  ColumnCacheMap column_cache;
  size_t max_groups_buffer_entry_guess = 1;

  { /* PerfectHashNoFallback (GroupByTest.cpp:100-152): SELECT COUNT(*) FROM t WHERE x = 1 GROUP BY str */
    RelAlgExecutionUnit u;
    const auto col = u.makeColumnVar(info.col_types[0], 0);
    u.simple_quals.push_back(u.makeBinOper(kEQ, col, u.makeConstant(int64_t(1))));
    u.groupby_exprs.push_back(u.makeColumnVar(info.col_types[3], 3));
    u.target_exprs.push_back(u.makeAggExpr(SQLTypeInfo(kINT, false), kCOUNT, -1));
    auto result = executor->executeWorkUnit(max_groups_buffer_entry_guess, true, {info}, u, CompilationOptions::defaults(ExecutorDeviceType::GPU),
                                            ExecutionOptions::defaults(), nullptr, false, column_cache);
    CHECK(result->rowCount() == size_t(1));
    auto row = result->getNextRow(false, false);
    CHECK(row.size() == size_t(1));
    CHECK(row[0].ival == 1 && !row[0].is_null);
    CHECK(result->getQueryMemDesc().query_desc_type == B2Q_GroupByPerfectHash);
  }
  { /* BaselineFallbackTest: sparse key; the first call must throw CardinalityEstimationRequired */
    RelAlgExecutionUnit u;
    u.groupby_exprs.push_back(u.makeColumnVar(info.col_types[2], 2));
    u.target_exprs.push_back(u.makeAggExpr(SQLTypeInfo(kINT, false), kCOUNT, -1));
    bool thrown = false;
    try {
      executor->executeWorkUnit(max_groups_buffer_entry_guess, true, {info}, u, CompilationOptions::defaults(), ExecutionOptions::defaults(), nullptr, false, column_cache);
    } catch (const CardinalityEstimationRequired&) { thrown = true; }
    CHECK(thrown);
    /* BaselineNoFilters: with an estimate => 2 rows, each COUNT 1 */
    max_groups_buffer_entry_guess = 4;
    auto result = executor->executeWorkUnit(max_groups_buffer_entry_guess, true, {info}, u, CompilationOptions::defaults(), ExecutionOptions::defaults(), nullptr, true, column_cache);
    CHECK(result->getQueryMemDesc().query_desc_type == B2Q_GroupByBaselineHash);
    CHECK(result->rowCount() == size_t(2));
    for (int i = 0; i < 2; ++i) { auto row = result->getNextRow(false, false); CHECK(row.size() == 1 && row[0].ival == 1); }
    CHECK(result->getNextRow(false, false).empty());
  }
  { /* a CPU device type must be refused: no CPU execution on this path */
    RelAlgExecutionUnit u;
    u.target_exprs.push_back(u.makeAggExpr(SQLTypeInfo(kINT, false), kCOUNT, -1));
    bool thrown = false;
    try {
      executor->executeWorkUnit(max_groups_buffer_entry_guess, true, {info}, u, CompilationOptions::defaults(ExecutorDeviceType::CPU), ExecutionOptions::defaults(), nullptr, false, column_cache);
    } catch (const QueryNotSupported&) { thrown = true; }
    CHECK(thrown);
  }
  { /* SELECT x, SUM(big), AVG(big) FROM t GROUP BY x */
    RelAlgExecutionUnit u;
    u.groupby_exprs.push_back(u.makeColumnVar(info.col_types[0], 0));
    u.target_exprs.push_back(u.makeColumnVar(info.col_types[0], 0));
    u.target_exprs.push_back(u.makeAggExpr(SQLTypeInfo(kBIGINT, false), kSUM, u.makeColumnVar(info.col_types[1], 1)));
    u.target_exprs.push_back(u.makeAggExpr(SQLTypeInfo(kDOUBLE, false), kAVG, u.makeColumnVar(info.col_types[1], 1)));
    auto result = executor->executeWorkUnit(max_groups_buffer_entry_guess, true, {info}, u, CompilationOptions::defaults(), ExecutionOptions::defaults(), nullptr, false, column_cache);
    CHECK(result->rowCount() == 2 && result->colCount() == 3);
    CHECK(result->getColType(2).get_type() == kDOUBLE);
    auto r0 = result->getNextRow(false, false);
    auto r1 = result->getNextRow(false, false);
    CHECK(r0[0].ival == 1 && r0[1].ival == 10 && r0[2].dval == 10.0);
    CHECK(r1[0].ival == 2 && r1[1].ival == 20 && r1[2].dval == 20.0);
  }
  { /* SELECT x, SUM(big) FROM t GROUP BY x ORDER BY 2 DESC LIMIT 1 — sort_info in the unit, and ResultSet::sort afterwards */
    RelAlgExecutionUnit u;
    u.groupby_exprs.push_back(u.makeColumnVar(info.col_types[0], 0));
    u.target_exprs.push_back(u.makeColumnVar(info.col_types[0], 0));
    u.target_exprs.push_back(u.makeAggExpr(SQLTypeInfo(kBIGINT, false), kSUM, u.makeColumnVar(info.col_types[1], 1)));
    u.sort_info.order_entries.push_back(B2QOrderEntry{2, 1, 1, {0, 0}});
    u.sort_info.limit = 1;
    auto result = executor->executeWorkUnit(max_groups_buffer_entry_guess, true, {info}, u, CompilationOptions::defaults(), ExecutionOptions::defaults(), nullptr, false, column_cache);
    CHECK(result->rowCount() == 1 && result->entryCount() == 1);
    auto r0 = result->getNextRow(false, false);
    CHECK(r0[0].ival == 2 && r0[1].ival == 20);
    CHECK(result->getNextRow(false, false).empty());
    u.sort_info = {};
    auto all = executor->executeWorkUnit(max_groups_buffer_entry_guess, true, {info}, u, CompilationOptions::defaults(), ExecutionOptions::defaults(), nullptr, false, column_cache);
    all->sort({B2QOrderEntry{1, 1, 0, {0, 0}}}, 0);
    CHECK(all->rowCount() == 2);
    CHECK(all->getNextRow(false, false)[0].ival == 2);
    CHECK(all->getNextRow(false, false)[0].ival == 1);
  }
  { /* SELECT d.attr, COUNT(*), SUM(t.big) FROM t JOIN d ON t.x = d.id GROUP BY d.attr — one INNER hash-join level */
    std::vector<int32_t> id{2, 1, 5}, attr{70, 70, 80};
    InputTableInfo dim;
    dim.col_types = {SQLTypeInfo(kINT, true), SQLTypeInfo(kINT, true)};
    dim.memory_level = MemoryLevel::CPU_LEVEL;
    FragmentInfo df;
    df.numTuples = 3;
    df.col_buffers = {id.data(), attr.data()};
    df.chunkStats.resize(2);
    df.chunkStats[0].int_min = 1; df.chunkStats[0].int_max = 5;
    df.chunkStats[1].int_min = 70; df.chunkStats[1].int_max = 80;
    dim.fragments.push_back(df);
    RelAlgExecutionUnit u;
    RelAlgExecutionUnit::JoinCondition jc;
    jc.quals.push_back(u.makeBinOper(kEQ, u.makeColumnVar(info.col_types[0], 0, 0), u.makeColumnVar(dim.col_types[0], 0, 1)));
    u.join_quals.push_back(jc);
    u.groupby_exprs.push_back(u.makeColumnVar(dim.col_types[1], 1, 1));
    u.target_exprs.push_back(u.makeColumnVar(dim.col_types[1], 1, 1));
    u.target_exprs.push_back(u.makeAggExpr(SQLTypeInfo(kINT, false), kCOUNT, -1));
    u.target_exprs.push_back(u.makeAggExpr(SQLTypeInfo(kBIGINT, false), kSUM, u.makeColumnVar(info.col_types[1], 1, 0)));
    auto result = executor->executeWorkUnit(max_groups_buffer_entry_guess, true, {info, dim}, u, CompilationOptions::defaults(), ExecutionOptions::defaults(), nullptr, false, column_cache);
    CHECK(result->rowCount() == 1);   /* both fact rows (x = 1, 2) map to attr 70 */
    auto r0 = result->getNextRow(false, false);
    CHECK(r0[0].ival == 70 && r0[1].ival == 2 && r0[2].ival == 30);
    CHECK(result->getQueryMemDesc().join_entry_count == 5);
  }
  { /* SELECT dd, COUNT(*), MAX(dd) FROM e WHERE dd >= DATE '1970-01-03' GROUP BY dd — DATE ENCODING DAYS(32) chunk
     * (SQLTypeInfo::get_compression() == kENCODING_DATE_IN_DAYS) with a $deleted$ column; row 3 is deleted */
    std::vector<int32_t> days{1, 2, 2, 4, INT32_MIN};
    std::vector<int8_t> deleted{0, 0, 0, 1, 0};
    InputTableInfo e;
    const SQLTypeInfo date_days(kDATE, false, kENCODING_DATE_IN_DAYS, 32);
    e.col_types = {date_days, SQLTypeInfo(kBOOLEAN, true)};
    e.memory_level = MemoryLevel::CPU_LEVEL;
    e.deleted_column = 1;
    FragmentInfo f;
    f.numTuples = 5;
    f.col_buffers = {days.data(), deleted.data()};
    f.chunkStats.resize(2);
    f.chunkStats[0].int_min = 1 * 86400; f.chunkStats[0].int_max = 4 * 86400; f.chunkStats[0].has_nulls = true; /* DateDaysEncoder keeps seconds */
    f.chunkStats[1].int_min = 0; f.chunkStats[1].int_max = 1;
    e.fragments.push_back(f);
    RelAlgExecutionUnit u;
    const ExprRef dd = u.makeColumnVar(date_days, 0, 0);
    u.simple_quals.push_back(u.makeBinOper(kGE, dd, u.makeConstant(SQLTypeInfo(kDATE, true), int64_t(2 * 86400))));
    u.groupby_exprs.push_back(dd);
    u.target_exprs.push_back(dd);
    u.target_exprs.push_back(u.makeAggExpr(SQLTypeInfo(kINT, false), kCOUNT, -1));
    u.target_exprs.push_back(u.makeAggExpr(date_days, kMAX, dd));
    auto result = executor->executeWorkUnit(max_groups_buffer_entry_guess, true, {e}, u, CompilationOptions::defaults(), ExecutionOptions::defaults(), nullptr, false, column_cache);
    CHECK(result->getQueryMemDesc().bucket == 86400 && result->getQueryMemDesc().entry_count == 4); /* days 2..4 + the NULL group */
    CHECK(result->rowCount() == 1);
    auto r0 = result->getNextRow(false, false);
    CHECK(r0[0].ival == 2 * 86400 && r0[1].ival == 2 && r0[2].ival == 2 * 86400);
    /* ColumnarResults of the same rows: DATE as int64, COUNT as int32 */
    ColumnarResults cols(*result, result->colCount(), {});
    CHECK(cols.size() == 1 && cols.getColumnBuffers().size() == 3);
    CHECK(*reinterpret_cast<const int64_t*>(cols.getColumnBuffers()[0]) == 2 * 86400);
    CHECK(*reinterpret_cast<const int32_t*>(cols.getColumnBuffers()[1]) == 2);
    CHECK(cols.getColumnType(1).get_type() == kINT && cols.getColumnType(0).get_type() == kDATE);
  }
  std::printf("boundary test ok\n");
  return 0;
}
