/* Compile-and-link check of the parts of include/b2q_executor.hpp the boundary test does not execute: the NDV estimator
 * unit, getNDVEstimator, the deleted-column compilation option, typed constants / UOper, ColumnarResults.  `main` builds
 * the objects but runs nothing on a device (this file is compiled and linked on the CPU-only box); what it does RUN is the
 * read-out half over hand-filled storage (resultSetFromStorage), which needs no device. */
#include <cstdio>

#include "b2q_executor.hpp"

using namespace b2q;

static std::shared_ptr<ResultSet> never_called(Executor& ex, const InputTableInfo& info) {
  RelAlgExecutionUnit u;
  const SQLTypeInfo date_days(kDATE, false, kENCODING_DATE_IN_DAYS, 32);
  const ExprRef d = u.makeColumnVar(date_days, 0, 0);
  u.quals.push_back(u.makeUOper(kNOT, u.makeUOper(kISNULL, d)));
  u.simple_quals.push_back(u.makeBinOper(kGE, d, u.makeConstant(SQLTypeInfo(kDATE, true), int64_t(86400))));
  RelAlgExecutionUnit::NDVEstimator est;
  est.expr_tuple.push_back(d);
  est.large = true;
  u.estimator = est;
  size_t guess = 0;
  ColumnCacheMap cache;
  CompilationOptions co = CompilationOptions::defaults();
  co.filter_on_deleted_column = false;
  auto rs = ex.executeWorkUnit(guess, true, {info}, u, co, ExecutionOptions::defaults(), nullptr, false, cache);
  std::printf("%zu\n", rs->getNDVEstimator());
  ColumnarResults cols(*rs, rs->colCount(), {}, true);
  std::printf("%zu %d\n", cols.size(), static_cast<int>(cols.getColumnType(0).get_type()));
  /* DECIMAL(10, 2): the constant's Datum is the scaled integer; getNextRow's decimal_to_double picks the read-out */
  const SQLTypeInfo dd(kDECIMAL, 10, 2, false);
  const ExprRef c = u.makeColumnVar(dd, 1, 0);
  u.quals.push_back(u.makeBinOper(kGT, c, u.makeConstant(SQLTypeInfo(kDECIMAL, 10, 2, true), int64_t(11100))));
  u.target_exprs.push_back(u.makeAggExpr(dd, kSUM, c));
  const auto row = rs->getNextRow(false, true);
  std::printf("%zu %d %d\n", row.size(), rs->getColType(0).get_scale(), static_cast<int>(dd.is_decimal()));
  return rs;
}

/* Runs on the CPU-only box: the read-out half through the C++ mirror, over hand-filled storage the way
 * Tests/ResultSetTest.cpp fills a ResultSetStorage.  SELECT x, COUNT(*), SUM(dd), AVG(dd) FROM t GROUP BY x with x INT NOT NULL
 * in [7, 8] and dd DECIMAL(10, 2): keyless perfect hash, five 8-byte slots per entry. */
static int readout_over_hand_filled_storage() {
  InputTableInfo info;
  const SQLTypeInfo x_ti(kINT, true), dd_ti(kDECIMAL, 10, 2, false);
  info.col_types = {x_ti, dd_ti};
  info.memory_level = MemoryLevel::CPU_LEVEL;
  FragmentInfo f;
  f.numTuples = 3;
  f.col_buffers = {nullptr, nullptr}; /* planning reads the chunk stats only */
  ChunkStats sx; sx.int_min = 7; sx.int_max = 8;
  ChunkStats sd; sd.int_min = 11110; sd.int_max = 22220; sd.has_nulls = true;
  f.chunkStats = {sx, sd};
  info.fragments.push_back(f);
  RelAlgExecutionUnit u;
  const ExprRef x = u.makeColumnVar(x_ti, 0), dd = u.makeColumnVar(dd_ti, 1);
  u.groupby_exprs.push_back(x);
  u.target_exprs = {x, u.makeAggExpr(SQLTypeInfo(kINT, false), kCOUNT, -1), u.makeAggExpr(dd_ti, kSUM, dd),
                    u.makeAggExpr(SQLTypeInfo(kDOUBLE, false), kAVG, dd)};
  Executor ex;
  const auto co = CompilationOptions::defaults();
  const auto eo = ExecutionOptions::defaults();
  /* rows: x = 7 twice (dd 111.10 and NULL), x = 8 once (222.20) */
  const int64_t storage[10] = {7, 2, 11110, 11110, 1, 8, 1, 22220, 22220, 1};
  auto rs = ex.resultSetFromStorage(reinterpret_cast<const int8_t*>(storage), sizeof(storage), 0, {info}, u, co, eo);
  const B2QPlan& qmd = rs->getQueryMemDesc();
  if (qmd.entry_count != 2 || !qmd.keyless_hash || qmd.row_size != 40 || rs->rowCount() != 2 || rs->colCount() != 4) return 1;
  if (rs->getColType(2).get_type() != kDECIMAL || rs->getColType(2).get_scale() != 2 || rs->getColType(3).get_type() != kDOUBLE) return 2;
  auto r0 = rs->getNextRow(false, true), r1 = rs->getNextRow(false, false);
  if (r0.size() != 4 || r0[0].ival != 7 || r0[1].ival != 2 || !r0[2].is_fp || r0[2].dval != 11110 / 100.0 || r0[3].dval != 11110 / (1 * 100.0)) return 3;
  if (r1.size() != 4 || r1[0].ival != 8 || r1[2].is_fp || r1[2].ival != 22220 || r1[3].dval != 222.2) return 4;
  if (!rs->getNextRow(false, true).empty()) return 5;
  /* an untouched entry (COUNT marker slot at its init value 0) does not exist */
  const int64_t one_empty[10] = {7, 2, 11110, 11110, 1, 0, 0, INT64_MIN, 0, 0};
  auto rs2 = ex.resultSetFromStorage(reinterpret_cast<const int8_t*>(one_empty), sizeof(one_empty), 0, {info}, u, co, eo);
  if (rs2->rowCount() != 1 || rs2->isRowAtEmpty(0) || !rs2->isRowAtEmpty(1)) return 6;
  /* getRowAt / getRowAtNoTranslations: random access by entry, an empty vector for an empty entry or past entryCount() */
  if (rs2->getRowAt(0).size() != rs2->colCount() || !rs2->getRowAtNoTranslations(1).empty() || !rs2->getRowAt(rs2->entryCount()).empty()) return 8;
  ColumnarResults cols(*rs2, 4, {});
  if (cols.size() != 1 || reinterpret_cast<const int64_t*>(cols.getColumnBuffers()[2])[0] != 11110) return 7;
  try { /* a buffer of another size is not this descriptor's */
    ex.resultSetFromStorage(reinterpret_cast<const int8_t*>(storage), 72, 0, {info}, u, co, eo);
    return 8;
  } catch (const QueryExecutionError&) {}
  return 0;
}

int main(int argc, char**) {
  const int rc = readout_over_hand_filled_storage();
  if (rc) { std::printf("read-out over hand-filled storage: check %d failed (%s)\n", rc, b2q_last_error_message()); return rc; }
  if (argc > 1000) { /* keeps the calls alive for the linker without ever making them */
    Executor ex;
    InputTableInfo info;
    never_called(ex, info);
  }
  std::printf("mirror surface links\n");
  return 0;
}
