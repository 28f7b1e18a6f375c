/*
 * qshost.h -- C entry points of libqshost.so: the C++ operator layer
 * (quickstep_b200/host/: RelationalOperator / WorkOrder subclasses with the
 * reference's names and contracts, QueryContext, Foreman/Worker scheduling,
 * storage blocks in the reference's physical layouts) driven as whole TPC-H
 * queries.  This is what tests/ and bench.py call to exercise the path the way
 * quickstep_cli_shell would: SQL text is replaced by the operator DAG the
 * reference's optimizer produces for the query, and
 * everything from QueryContext construction to the result rows runs in C++
 * over the libqsgpu.so C ABI.  No torch, no Python objects in the signatures.
 */
#ifndef QSHOST_H_
#define QSHOST_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qshost_db *qshost_db_t;

/* Relations of the TPC-H subset (attribute order = quickstep_b200/tpch.py):
 *   CUSTOMER  c_custkey INT, c_mktsegment CHAR(10)
 *   ORDERS    o_orderkey INT, o_custkey INT, o_orderdate DATE, o_shippriority INT
 *   LINEITEM  l_orderkey INT, l_quantity, l_extendedprice, l_discount, l_tax DOUBLE,
 *             l_returnflag CHAR(1), l_linestatus CHAR(1), l_shipdate DATE            */
enum { QSHOST_CUSTOMER = 0, QSHOST_ORDERS = 1, QSHOST_LINEITEM = 2 };
/* storage/StorageBlockLayout.proto tuple store types the path reads */
enum { QSHOST_BASIC_COLUMN_STORE = 0, QSHOST_COMPRESSED_COLUMN_STORE = 1, QSHOST_SPLIT_ROW_STORE = 2 };

/* dev: CUDA device (qsgpu_init is called for it); num_workers: Worker threads. */
int qshost_db_create(int dev, int num_workers, qshost_db_t *out);
int qshost_db_destroy(qshost_db_t db);
/* Several GPUs, one process per GPU: `comm` is a qsgpu_comm_t (include/qsgpu.h) of this process's device.  Set it
 * BEFORE loading: every relation loaded afterwards is this rank's PARTITION of the relation (lineitem block-partitioned
 * on l_orderkey boundaries, orders and customer in shares), statistics and row counts are reduced over the ranks,
 * and the query entry points return the COMPLETE answer on every rank: partial aggregation states are merged
 * (Q1, Q6), LIP filters OR-reduced, the filtered orders all-gathered for a broadcast join and the per-rank top-k
 * candidates gathered (Q3) -- all inside the C++ operator layer, through the collectives of the C ABI. */
int qshost_db_set_comm(qshost_db_t db, void *comm);
/* Q3's join with several GPUs.  0 (default): partition-wise -- orders and lineitem are partitioned on the order key by the
 * same scheme (the caller loads, on every rank, the lineitem rows of exactly the orders it loads there), so every GPU
 * builds and probes the hash table of its own partition (the reference's per-partition hash tables,
 * query_execution/QueryContext.cpp:78-97).  1: broadcast -- the filtered orders of all ranks are all-gathered and every
 * GPU builds the whole table (what a build side partitioned any other way needs). */
int qshost_db_set_join_mode(qshost_db_t db, int mode);
/* COPY ... + \analyze: cut native-width columns into storage blocks of rows_per_block
 * tuples in `layout`, and record min/max of the key attributes.  May be called again
 * for the same relation to replace it. */
int qshost_db_load(qshost_db_t db, int which, const void *const *columns, uint64_t n_rows,
                   uint64_t rows_per_block, int layout);
/* Drop the HBM image of a relation: the next query stages its blocks again (cold run). */
int qshost_db_evict(qshost_db_t db, int which);
/* Keep dictionary-compressed attributes of the base relations as codes in HBM (qsgpu_relation_set_dictionary) instead
 * of decoding them to native columns while staging: scans then run on the codes
 * (CompressedTupleStorageSubBlock::getMatchesForPredicate, storage/CompressedTupleStorageSubBlock.cpp:160-251).  Drops
 * the HBM images so that the next query stages in the new form.  Off by default. */
int qshost_db_set_code_resident(qshost_db_t db, int on);
/* Code width (0 = native) and dictionary entries attribute `attr` of a relation's current HBM image uses. */
int qshost_db_resident_coding(qshost_db_t db, int which, uint32_t attr, uint32_t *code_width, uint32_t *n_entries);
/* Bytes of the relation's block images on the host / number of blocks. */
int qshost_db_stats(qshost_db_t db, int which, uint64_t *host_bytes, uint64_t *n_blocks, uint64_t *n_rows);
/* Cap on the rows one GPU work order covers (0 = one work order per run of blocks). */
int qshost_set_rows_per_workorder(uint64_t rows);

typedef struct qshost_q1_row {
  char l_returnflag, l_linestatus;
  char pad[6];
  double sum_qty, sum_base_price, sum_disc_price, sum_charge, avg_qty, avg_price, avg_disc;
  int64_t count_order;
  double sum_disc;   /* SUM(l_discount): not in the query's select list; kept so that partial results of
                        lineitem partitions (one per GPU) can be merged exactly (AVG = SUM / COUNT) */
} qshost_q1_row;

typedef struct qshost_q3_row {
  int32_t l_orderkey;
  int32_t o_shippriority;
  double revenue;
  int32_t year;
  uint8_t month, day;
  uint8_t pad[2];
} qshost_q3_row;

/* benchmarks/tpch/queries/01.sql, 06.sql, 03.sql.  *n_rows: in = capacity of rows[], out = rows
 * returned.  *work_orders (optional): work orders the scheduler executed for the query. */
int qshost_q1(qshost_db_t db, qshost_q1_row *rows, uint32_t *n_rows, uint64_t *work_orders);
int qshost_q6(qshost_db_t db, double *revenue, int *is_null, uint64_t *work_orders);
int qshost_q3(qshost_db_t db, qshost_q3_row *rows, uint32_t *n_rows, uint64_t *work_orders);

/* The result relation of the last query as the InsertDestination wrote it on the host: block `index` (0, 1, ...) of
 * SplitRowStore blocks in the reference's own layout ([int header_len][StorageBlockHeader proto][sub-block header]
 * [occupancy bitmap][tuple slots], storage/SplitRowStoreTupleStorageSubBlock.cpp:103-179), valid until the next query.
 * Returns non-zero when there is no such block. */
int qshost_result_block(qshost_db_t db, uint32_t index, const void **memory, uint64_t *bytes, uint64_t *n_tuples);

/* Per-operator profile of the last query: one text line per operator (index, name, work orders, ms inside
 * execute(), ms inside getAllWorkOrders()); the analogue of -profile_and_report_workorder_perf. */
int qshost_last_profile(qshost_db_t db, char *buf, uint64_t buf_bytes);

#ifdef __cplusplus
}
#endif
#endif  /* QSHOST_H_ */
