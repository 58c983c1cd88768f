// db_query — drop-in CLI of the reference's plain ADC query tool (db_query.cpp:48-134) on top of
// the B200 float scan:  db_query [-r R] [-m MA] [-b BATCH_SIZE] [-g GPU] [-o results.bin]
//                                db_file query_file groundtruth_file
// Same defaults (r=100, ma=1; batch=1 -> here "all queries in one batch") and the same CSV on
// stdout.  Quantisers: every pair of get_scan_func (query_common.hpp:122-147): (16,4) (32,4) (4,8) (8,8) (16,8)
// (2,16) (4,16) (8,16).  -o dumps per query r uint32 ids then r float32
// distances, ascending by distance, for tests.
#include <unistd.h>

#include <cstdio>

#include "query_common.hpp"

struct cmdargs : query_args {
    int batch_size;
    int gpu;
    const char* out_file;
};

static void usage() {
    std::cerr << "Usage: db_query: [-r R] [-m MA] [-b BATCH_SIZE] [-g GPU] [-o results.bin] "
              << "[db_file] [query_file] [groundtruth_file]" << std::endl;
    std::exit(1);
}

static void parse_args(cmdargs& args, int argc, char* argv[]) {
    int opt;
    args.ma = 1;
    args.r = 100;
    args.batch_size = 1;
    args.gpu = 0;
    args.out_file = nullptr;
    while ((opt = getopt(argc, argv, "r:m:b:g:o:")) != -1) {
        switch (opt) {
        case 'r': args.r = std::atoi(optarg); break;
        case 'm': args.ma = std::atoi(optarg); break;
        case 'b': args.batch_size = std::atoi(optarg); break;
        case 'g': args.gpu = std::atoi(optarg); break;
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
    std::cerr << "Database file: " << args.db_file << std::endl;
    std::unique_ptr<base_db> db = load_database(args.db_file);
    db->print(std::cerr);
    std::cerr << std::endl;
    query_metrics total_metrics;
    double total_recall = 0;
    std::unique_ptr<scanner_gpu_simple> scanner(new scanner_gpu_simple(args.gpu));
    // the reference switches engines on -b (db_query.cpp:97-113); the GPU engine is always batched
    nns_engine_gpu_adc engine(std::move(scanner), *db, args.ma, args.r, args.batch_size == 1 ? 0 : args.batch_size);
    std::vector<unsigned> keys;
    std::vector<float> vals;
    process_queries<nns_engine_gpu_adc, scanner_gpu_simple::BhType, query_metrics, float>(
        args, *db, engine, total_metrics, total_recall, args.out_file ? &keys : nullptr, args.out_file ? &vals : nullptr);
    std::cout << "r,recall,ma,adc_type," << query_metrics::header_string << std::endl;   // db_query.cpp:116-119
    std::cout << args.r << "," << total_recall << "," << args.ma << ",adc," << total_metrics << std::endl;
    if (args.out_file) {
        FILE* f = std::fopen(args.out_file, "wb");
        if (!f) { std::cerr << "Could not write " << args.out_file << std::endl; return 1; }
        const size_t nq = keys.size() / args.r;
        for (size_t q = 0; q < nq; ++q) {
            std::fwrite(keys.data() + q * args.r, 4, args.r, f);
            std::fwrite(vals.data() + q * args.r, 4, args.r, f);
        }
        std::fclose(f);
    }
    return 0;
}
