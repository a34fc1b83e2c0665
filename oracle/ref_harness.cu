// oracle/ref_harness.cu — TEST INFRASTRUCTURE: runs the UNMODIFIED reference XMtrustregion
// (/root/reference/XM/include/XM/trustregion.h:77-724, included read-only where it lies) so that
//   (1) the NumPy oracle and the CUDA product path can be pinned against real reference output, and
//   (2) bench.py --impl reference can time the reference's own cuBLAS code path on the same B200.
// The driver logic below is a small re-statement of what XM_main.cu:312-401 (solve_rank3) does around the
// call, generalised to an arbitrary start point / rank / line-search step so rank>3 calls can be replayed.
// Nothing here is shipped; the product never links it.  Build: oracle/Makefile -> oracle/_ref/xm_ref_harness.
//
// usage: xm_ref_harness <dir> <rank> <tol> <lam> <max_time> [ls_step] [repeat]
//   reads  <dir>/Q.bin ; optional <dir>/R_ini.bin (3N x rank) , <dir>/s_ini.bin (N x 1), <dir>/v_ini.bin (3N x 1)
//   writes <dir>/R_ref.bin , <dir>/s_ref.bin ; prints the reference's own per-iteration table and one
//   line  "REFJSON {...}"  with timings (ms) of [H2D of Q], [XMtrustregion], [D2H] for each repeat.
#include <XM/trustregion.h>
#include <string>
#include <cstdio>
#include <cstdlib>

static bool read_bin(const std::string& fn, std::vector<double>& m, int& rows, int& cols) {
    FILE* f = fopen(fn.c_str(), "rb");
    if (!f) return false;
    if (fread(&rows, 4, 1, f) != 1 || fread(&cols, 4, 1, f) != 1) { fclose(f); return false; }
    m.resize((size_t)rows * cols);
    size_t got = fread(m.data(), sizeof(double), m.size(), f);
    fclose(f);
    return got == m.size();
}
static void write_bin(const std::string& fn, const std::vector<double>& m, int rows, int cols) {
    FILE* f = fopen(fn.c_str(), "wb");
    fwrite(&rows, 4, 1, f); fwrite(&cols, 4, 1, f);
    fwrite(m.data(), sizeof(double), m.size(), f);
    fclose(f);
}

int main(int argc, char** argv) {
    if (argc < 6) { fprintf(stderr, "usage: %s dir rank tol lam max_time [ls_step] [repeat]\n", argv[0]); return 2; }
    std::string dir = argv[1];
    size_s o = (size_s)atoi(argv[2]);
    double tol = atof(argv[3]), lam = atof(argv[4]), max_time = atof(argv[5]);
    double ls_step = argc > 6 ? atof(argv[6]) : 0.0;
    int repeat = argc > 7 ? atoi(argv[7]) : 1;

    std::vector<double> Q_h; int rows = 0, cols = 0;
    if (!read_bin(dir + "/Q.bin", Q_h, rows, cols)) { fprintf(stderr, "cannot read Q.bin\n"); return 3; }
    size_s n = rows / 3;

    std::vector<double> R0_h((size_t)3 * n * o, 0.0), s0_h(n, 1.0), v_h((size_t)3 * n, 0.0), tmp;
    int rr, cc;
    if (read_bin(dir + "/R_ini.bin", tmp, rr, cc) && (size_t)rr * cc == R0_h.size()) R0_h = tmp;
    else for (size_l i = 0; i < n; ++i) { R0_h[3*i] = 1.0; R0_h[3*i + 3*n + 1] = 1.0; R0_h[3*i + 6*n + 2] = 1.0; }
    if (read_bin(dir + "/s_ini.bin", tmp, rr, cc) && (size_t)rr * cc == n) s0_h = tmp;
    if (read_bin(dir + "/v_ini.bin", tmp, rr, cc) && (size_t)rr * cc == (size_t)3 * n) v_h = tmp;

    std::string json = "REFJSON {\"n\": " + std::to_string(n) + ", \"rank\": " + std::to_string(o) + ", \"runs\": [";
    std::vector<double> R_h((size_t)3 * n * o), s_h(n);
    for (int rep = 0; rep < repeat; ++rep) {
        cudaEvent_t e0, e1, e2, e3;
        cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
        cudaDeviceSynchronize();
        auto w0 = std::chrono::high_resolution_clock::now();
        cudaEventRecord(e0);
        opt_var C({3*n, 3*n});
        C.SynchronizeHostToDevice(Q_h.data());
        cudaEventRecord(e1);
        opt_var v({3*n});
        v.SynchronizeHostToDevice(v_h.data());
        opt_var R0({3*n, o});
        R0.SynchronizeHostToDevice(R0_h.data());
        opt_var s0_ex({n});
        s0_ex.SynchronizeHostToDevice(s0_h.data());
        opt_var s0; s0.vals = s0_ex.vals + 1; s0.num_dims = 1; s0.dimensions = new size_s[1];
        s0.dimensions[0] = n - 1; s0.total_size = n - 1;
        opt_var s_ex({n});
        s_ex.SynchronizeHostToDevice(s0_h.data());
        opt_var s; s.vals = s_ex.vals + 1; s.num_dims = 1; s.dimensions = new size_s[1];
        s.dimensions[0] = n - 1; s.total_size = n - 1;
        opt_var R({3*n, o});
        double gradtol = tol, primal = 0;
        cudaDeviceSynchronize();
        auto t0 = std::chrono::high_resolution_clock::now();
        XMtrustregion(C, R0, s0, R, s, lam, gradtol, ls_step, v, &primal, max_time);
        cudaDeviceSynchronize();
        auto t1 = std::chrono::high_resolution_clock::now();
        cudaEventRecord(e2);
        cudaMemcpy(R_h.data(), R.vals, sizeof(double) * R_h.size(), cudaMemcpyDeviceToHost);
        cudaMemcpy(s_h.data(), s_ex.vals, sizeof(double) * n, cudaMemcpyDeviceToHost);
        cudaEventRecord(e3);
        cudaEventSynchronize(e3);
        auto w1 = std::chrono::high_resolution_clock::now();
        float h2d = 0, d2h = 0;
        cudaEventElapsedTime(&h2d, e0, e1); cudaEventElapsedTime(&d2h, e2, e3);
        double tr_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        double wall_ms = std::chrono::duration<double, std::milli>(w1 - w0).count();
        char buf[512];
        snprintf(buf, sizeof buf, "%s{\"h2d_q_ms\": %.4f, \"tr_ms\": %.4f, \"d2h_ms\": %.4f, \"wall_ms\": %.4f, \"primal\": %.17g, \"gradtol_out\": %.3e}",
                 rep ? ", " : "", h2d, tr_ms, d2h, wall_ms, primal, gradtol);
        json += buf;
        s0.vals = nullptr; s.vals = nullptr;
        fflush(stdout);
    }
    json += "]}";
    write_bin(dir + "/R_ref.bin", R_h, 3 * n, o);
    write_bin(dir + "/s_ref.bin", s_h, n, 1);
    printf("%s\n", json.c_str());
    return 0;
}
