// seaweedfs_b200/csrc/codegen_main.cc — build-time front end of codegen.cc.
//   swec_codegen --rs K M            parity rows of the RS(K,M) generator (the encode matrix)
//   swec_codegen --rows R K c…       an explicit R×K matrix (row-major decimal bytes)
//   options: --no-basis --no-cse --name StructName
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "codegen.h"

int main(int argc, char** argv) {
    swec::CodegenOptions opt;
    swec::Matrix rows;
    std::string name = "SwecGen";
    for (int a = 1; a < argc; a++) {
        if (!strcmp(argv[a], "--no-basis")) opt.optimise_basis = false;
        else if (!strcmp(argv[a], "--no-cse")) opt.extract_common = false;
        else if (!strcmp(argv[a], "--name") && a + 1 < argc) name = argv[++a];
        else if (!strcmp(argv[a], "--rs") && a + 2 < argc) {
            const int k = atoi(argv[a + 1]), m = atoi(argv[a + 2]);
            a += 2;
            swec::Matrix gen = swec::rs_generator(k, m);
            rows = swec::Matrix(m, k);
            for (int p = 0; p < m; p++)
                for (int i = 0; i < k; i++) rows.at(p, i) = gen.at(k + p, i);
        } else if (!strcmp(argv[a], "--rows") && a + 2 < argc) {
            const int r = atoi(argv[a + 1]), k = atoi(argv[a + 2]);
            a += 2;
            rows = swec::Matrix(r, k);
            for (int i = 0; i < r * k && a + 1 < argc; i++) rows.v[size_t(i)] = uint8_t(atoi(argv[++a]));
        } else {
            fprintf(stderr, "usage: %s [--no-basis] [--no-cse] (--rs K M | --rows R K c...)\n", argv[0]);
            return 2;
        }
    }
    if (rows.rows == 0) return 2;
    swec::CodegenStats st;
    fputs(swec::generate_combine(rows, name, opt, &st).c_str(), stdout);
    return 0;
}
