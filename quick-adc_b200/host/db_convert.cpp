// db_convert — rewrites a database file in the other format (no GPU needed):
//     db_convert in_db out_db
// in_db may be a database written by the reference's tools (cereal archive of flat_db/index_db)
// or a .qdb container; out_db ending in ".qdb" is written as a container, any other name in the
// reference's archive layout (see host/databases.hpp).
#include "databases.hpp"

int main(int argc, char* argv[]) {
    if (argc != 3) {
        std::cerr << "Usage: db_convert in_db out_db" << std::endl;
        return 1;
    }
    std::unique_ptr<base_db> db = load_database(argv[1]);
    db->print(std::cerr);
    std::cerr << std::endl;
    unsigned long total = 0;
    for (int p = 0; p < db->partition_count(); ++p) {
        const std::uint8_t* codes;
        unsigned* labels;
        unsigned size;
        db->get_partition(p, codes, labels, size);
        total += size;
    }
    std::cerr << "Vectors: " << total << std::endl;
    if (!save_database(*db, argv[2])) {
        std::cerr << "Could not write " << argv[2] << std::endl;
        return 1;
    }
    return 0;
}
