// spada-sim -- CLI drop-in for the reference binary (main.rs:30-121), compiled host side.
//
//   spada-sim <simulator> <accelerator> <category> <workload> <configuration> [-p]
//
// Same argv grammar, same loaders' behaviour, same stdout sections; C = A x B runs on the B200
// through include/spada_b200.h.  The simulated-hardware counters are replaced by analytic element
// counts (simulator.rs:1008-1032 are timing-model outputs, out of scope).  Only AccurateSimu is
// implemented, like upstream (main.rs:119 panics for the other modes).  Failures end the process
// with a non-zero status and a message on stderr, the analogue of the reference's unwrap()/panic!.
#include <cstdio>
#include <exception>

#include "spada_host.hpp"

using namespace spada_host;

static int run(int argc, char** argv) {
    Cli cli = parse_args(argc, argv);
    OmegaConfig cfg = parse_config(cli.configuration);
    GEMM gemm = (cli.category == "NN") ? load_pickled_gemms(cfg.nn_filepath, cli.workload)
                                       : GEMM::from_mat(cli.workload, load_mm_mat(cfg.ss_filepath, cli.workload));
    const CsrMat& B = gemm.B();
    if (gemm.a.rows == 0 || B.rows == 0) throw std::runtime_error("attempt to divide by zero");  // main.rs:44-45
    size_t a_avg = gemm.a.nnz() / gemm.a.rows, b_avg = B.nnz() / B.rows;
    std::printf("Get GEMM %s\n", gemm.name.c_str());
    std::printf("%s\n", gemm.display().c_str());
    std::printf("Avg row len of A: %zu, Avg row len of B: %zu\n", a_avg, b_avg);
    std::fflush(stdout);
    if (cli.simulator != "AccurateSimu") throw std::runtime_error("Unimplemented simulator " + cli.simulator);

    if (const char* dump = std::getenv("SPADA_B200_DUMP_OPERANDS")) {  // test hook: loader parity vs scipy
        FILE* f = std::fopen(dump, "wb");
        const CsrMat* mats[2] = {&gemm.a, &B};
        for (const CsrMat* m : mats) {
            uint64_t hdr[3] = {m->rows, m->cols, m->nnz()};
            std::fwrite(hdr, 8, 3, f);
            std::fwrite(m->indptr.data(), 8, m->indptr.size(), f);
            std::fwrite(m->indices.data(), 8, m->indices.size(), f);
            std::fwrite(m->data.data(), 8, m->data.size(), f);
        }
        std::fclose(f);
    }

    auto storages = CsrMatStorage::init_with_gemm(gemm);
    CsrMatStorage &dram_a = storages.first, &dram_b = storages.second;
    if (cli.preprocess) dram_a.reorder_row(sort_by_length(dram_a));  // never changes C (simulator.rs:1039-1060)
    size_t output_base_addr = dram_b.indptr.size();
    std::array<size_t, 2> block_shape = cfg.block_shape;
    if (cli.accelerator == "Op") block_shape = {cfg.lane_num, 1};  // main.rs:67-72

    Simulator sim(cfg.pe_num, cfg.at_num, cfg.lane_num, cfg.cache_size, cfg.word_byte, output_base_addr, block_shape,
                  dram_a, dram_b, cli.accelerator, cfg.mem_latency, cfg.cache_latency, cfg.freq, cfg.channel,
                  cfg.bandwidth_per_channel);
    sim.execute();
    std::vector<CsrRow> result = sim.get_exec_result();
    if (const char* sink = std::getenv("SPADA_B200_DUMP_C"))   // result sink: all of C + digest (stderr keeps stdout = the reference's)
        std::fprintf(stderr, "%s\n", dump_result(sink, result, dram_b.mat_shape[0]).c_str());
    auto a_count = sim.get_a_mat_stat(), b_count = sim.get_b_mat_stat(), c_count = sim.get_c_mat_stat();
    auto cache_count = sim.get_cache_stat();
    std::printf("-----Result-----\n");
    std::printf("-----Access count\n");
    std::printf("Execution count: %zu\n", sim.get_exec_cycle());
    std::printf("A matrix count: read %zu write %zu\n", a_count[0], a_count[1]);
    std::printf("B matrix count: read %zu write %zu\n", b_count[0], b_count[1]);
    std::printf("C matrix count: read %zu write %zu\n", c_count[0], c_count[1]);
    std::printf("Cache count: read %zu write %zu\n", cache_count[0], cache_count[1]);
    std::printf("-----Output product matrix\n");
    for (size_t i = 0; i < std::min<size_t>(result.size(), 10); ++i) std::printf("%s\n", result[i].display().c_str());
    return 0;
}

int main(int argc, char** argv) {
    try {
        return run(argc, argv);
    } catch (const std::invalid_argument& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 2;  // clap's exit status for usage errors
    } catch (const std::exception& e) {
        std::fflush(stdout);
        std::fprintf(stderr, "thread 'main' panicked at '%s'\n", e.what());
        return 101;  // Rust's exit status for a panic
    }
}
