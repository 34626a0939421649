/* Compile-and-link check of the parts of include/b2q_executor.hpp the boundary test does not execute: the NDV estimator
 * unit, getNDVEstimator, the deleted-column compilation option, typed constants / UOper, ColumnarResults.  `main` builds
 * the objects but runs nothing on a device (this file is compiled and linked on the CPU-only box). */
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

int main(int argc, char**) {
  if (argc > 1000) { /* keeps the calls alive for the linker without ever making them */
    Executor ex;
    InputTableInfo info;
    never_called(ex, info);
  }
  std::printf("mirror surface links\n");
  return 0;
}
