// db_query_4 — drop-in CLI of the reference's Quick ADC query tool (db_query_4.cpp:312-414) on
// top of the B200 path:  db_query_4 [-r R] [-m MA] [-k KEEP_PERCENT] [-b BATCH_SIZE] [-g GPU[,GPU...]]
//                                   [-o results.bin] db_file query_file groundtruth_file
// -g takes one device ordinal or a comma-separated list (-g 0,1,2,3,4,5,6,7): the database is then sharded over
// those GPUs by this one process (qadc_multi_*: NCCL all-gather of the per-GPU top-r lists + merge).
// Same defaults (r=100, ma=1, keep=1 %, batch=1 -> here "all queries in one batch", since the
// GPU engine is always batched) and the same CSV on stdout.  -o dumps the per-query results
// (r uint32 ids then r int8 distances per query, ascending by distance) for tests.
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <strings.h>

#include "query_common.hpp"

struct cmdargs : query_args {
    float keep;
    int batch_size;
    std::vector<int> gpus;
    const char* out_file;
};

static void usage() {
    std::cerr << "Usage: db_query_4 [-r R] [-m MA] [-k KEEP_PERCENT] [-b BATCH_SIZE] [-g GPU[,GPU...]] [-o results.bin] "
              << "[db_file] [query_file] [groundtruth_file]" << std::endl;
    std::exit(1);
}

static void parse_args(cmdargs& args, int argc, char* argv[]) {
    const float ONE_PERCENT = 0.01;
    int opt;
    args.ma = 1;
    args.r = 100;
    args.keep = 1 * ONE_PERCENT;
    args.batch_size = 1;
    args.gpus.assign(1, 0);
    args.out_file = nullptr;
    while ((opt = getopt(argc, argv, "r:m:b:k:g:o:")) != -1) {
        switch (opt) {
        case 'r': args.r = std::atoi(optarg); break;
        case 'm': args.ma = std::atoi(optarg); break;
        case 'b': args.batch_size = std::atoi(optarg); break;
        case 'k': args.keep = std::atof(optarg) * ONE_PERCENT; break;
        case 'g': {
            args.gpus.clear();
            for (const char* p = optarg; *p;) {
                char* end;
                args.gpus.push_back(static_cast<int>(std::strtol(p, &end, 10)));
                if (end == p) usage();
                p = (*end == ',') ? end + 1 : end;
                if (*end && *end != ',') usage();
            }
            if (args.gpus.empty()) usage();
            break;
        }
        case 'o': args.out_file = optarg; break;
        default: usage();
        }
    }
    if (argc - optind < 3) usage();
    args.db_file = argv[optind];
    args.query_file = argv[optind + 1];
    args.groundtruth_file = argv[optind + 2];
}

int main(int argc, char* argv[]) {
    cmdargs args;
    parse_args(args, argc, argv);
    // stdout carries the reference's CSV and nothing else: NCCL's own banner / debug lines (NCCL_DEBUG) go to stderr
    // (NCCL honours NCCL_DEBUG_FILE only above the VERSION level, so a VERSION request is raised to WARN)
    if (args.gpus.size() > 1) {
        setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
        const char* lvl = std::getenv("NCCL_DEBUG");
        if (lvl && strcasecmp(lvl, "VERSION") == 0) setenv("NCCL_DEBUG", "WARN", 1);
    }
    std::cerr << "Database file: " << args.db_file << std::endl;
    std::unique_ptr<base_db> db = load_database(args.db_file);
    if (db->pq->sq_bits != 4) {
        std::cerr << "Quantizer must have  sq_bits=4" << std::endl;   // load_database_check, db_query_4.cpp:393-402
        return 1;
    }
    query_metrics total_metrics;
    double total_recall = 0;
    std::unique_ptr<scanner_gpu_4> scanner(new scanner_gpu_4(args.keep, args.gpus));
    nns_engine_gpu engine(std::move(scanner), *db, args.ma, args.r, args.batch_size == 1 ? -1 : args.batch_size);
    std::vector<unsigned> keys;
    std::vector<std::int8_t> vals;
    process_queries<nns_engine_gpu, scanner_gpu_4::BhType>(args, *db, engine, total_metrics, total_recall,
                                                            args.out_file ? &keys : nullptr, args.out_file ? &vals : nullptr);
    std::cout << "r,recall,ma,adc_type,keep," << query_metrics::header_string << std::endl;
    std::cout << args.r << "," << total_recall << "," << args.ma << "," << "qadc," << args.keep << "," << total_metrics
              << std::endl;
    if (args.out_file) {
        FILE* f = std::fopen(args.out_file, "wb");
        if (!f) { std::cerr << "Could not write " << args.out_file << std::endl; return 1; }
        const size_t nq = keys.size() / args.r;
        for (size_t q = 0; q < nq; ++q) {
            std::fwrite(keys.data() + q * args.r, 4, args.r, f);
            std::fwrite(vals.data() + q * args.r, 1, args.r, f);
        }
        std::fclose(f);
    }
    return 0;
}
