// db_build — creates a database file from a quantiser and a base-vector file, the job of the
// reference's flatdb_create / indexdb_create2 + db_add chain (flatdb_create.cpp, db_add.cpp) with
// the encoding done on the GPU:
//     db_build [-c coarse_centroids.fvecs] [-g GPU] quantizer.(o)pq.data base.(f|b)vecs out_db
// out_db ending in ".qdb" is written as a .qdb container, any other name in the reference's own
// archive layout (host/databases.hpp), which the reference's tools read.
// Without -c a flat database is written, with -c an inverted-list database over those centroids
// (training the centroids — k-means, indexdb_create1.cpp — is out of scope).  Vectors are added
// in chunks of one million like db_add (db_add.cpp:52-77), ids = position in the base file.
#include <unistd.h>

#include "databases.hpp"
#include "vector_io.hpp"

int main(int argc, char* argv[]) {
    const char* coarse = nullptr;
    int gpu = 0, opt;
    while ((opt = getopt(argc, argv, "c:g:")) != -1) {
        if (opt == 'c') coarse = optarg;
        else if (opt == 'g') gpu = std::atoi(optarg);
        else { std::cerr << "Usage: db_build [-c centroids.fvecs] [-g GPU] pq_file base_file out_db" << std::endl; return 1; }
    }
    if (argc - optind < 3) { std::cerr << "Usage: db_build [-c centroids.fvecs] [-g GPU] pq_file base_file out_db" << std::endl; return 1; }
    std::unique_ptr<base_pq> pq = pq_from_data_file(argv[optind]);
    vectors_owner<float> base = load_vectors_by_extension(argv[optind + 1]);
    if (base.dimension != pq->dim) { std::cerr << "Base dimension " << base.dimension << " != quantizer dimension " << pq->dim << std::endl; return 1; }
    qadc_ctx* enc = nullptr;
    if (qadc_create(gpu, nullptr, &enc)) { std::cerr << "qadc_create: " << qadc_last_error(nullptr) << std::endl; return 1; }
    if (qadc_set_pq(enc, pq->dim, pq->sq_count, pq->sq_bits, pq->centroids_flat.data(), pq->rotation_ptr())) {
        std::cerr << "qadc_set_pq: " << qadc_last_error(enc) << std::endl; return 1;
    }
    std::unique_ptr<base_db> db;
    if (coarse) {
        vectors_owner<float> cents = load_vectors_by_extension(coarse);
        if (cents.dimension != pq->dim) { std::cerr << "Centroid dimension mismatch" << std::endl; return 1; }
        auto* x = new index_db;
        db.reset(x);
        x->part_count = static_cast<int>(cents.count);
        x->centroids.assign(cents.get(0), cents.get(0) + cents.count * cents.dimension);
        x->partitions.resize(x->part_count);
        x->labels.resize(x->part_count);
        if (qadc_set_coarse(enc, x->part_count, x->centroids.data())) { std::cerr << "qadc_set_coarse: " << qadc_last_error(enc) << std::endl; return 1; }
    } else {
        db.reset(new flat_db);
    }
    db->pq = std::move(pq);
    const long chunk = 1000000;
    for (long off = 0; off < base.count; off += chunk) {
        const unsigned n = static_cast<unsigned>(std::min(chunk, base.count - off));
        db->add_vectors(enc, base.get(off), n, static_cast<unsigned>(off));
        std::cerr << "Added " << off + n << "/" << base.count << "\r";
    }
    std::cerr << std::endl;
    db->print(std::cerr);
    std::cerr << std::endl;
    if (!save_database(*db, argv[optind + 2])) { std::cerr << "Could not write " << argv[optind + 2] << std::endl; return 1; }
    qadc_destroy(enc);
    return 0;
}
