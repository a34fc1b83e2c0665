// XM_main.cpp — host side of the pybind11 module "XM" (same module name, function names, positional arguments,
// file formats and stdout banners as the reference's XM/src/XM_main.cu:35-408), written against the C-ABI of
// libxm_b200.so (include/xm_b200.h).  All device work — Q.Y, Riemannian gradient / Hessian-vector, retraction,
// the whole trust-region loop, the certificate — happens behind that ABI in hand-written sm_100a kernels.
//
//   XM.solve(dataset_path, max_rank, tol, lam, max_time)          -> None   reads <path>/Q.bin, writes R.bin, s.bin
//   XM.solve_rank3(dataset_path, max_rank, tol, lam, max_time)    -> None   one rank-3 solve, no certificate
//   XM.solve_rebuttle(dataset_path, max_rank, tol, lam, max_time) -> int    {1 certified, 2 max rank, -2 line search}
#include <pybind11/pybind11.h>

#include <cstdio>
#include <algorithm>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "xm_b200.h"

namespace py = pybind11;

namespace {

bool load_bin(const std::string& fn, std::vector<double>& m, int& rows, int& cols) {
    std::ifstream f(fn, std::ios::binary);
    if (!f) { std::cerr << "cannot open file" << std::endl; return false; }   // XM_main.cu:21-24
    f.read(reinterpret_cast<char*>(&rows), sizeof(int));
    f.read(reinterpret_cast<char*>(&cols), sizeof(int));
    m.resize((size_t)rows * cols);
    f.read(reinterpret_cast<char*>(m.data()), sizeof(double) * m.size());
    printf("rows: %d, cols: %d\n", rows, cols);
    return true;
}

void save_bin(const std::string& fn, const double* m, int rows, int cols) {
    std::ofstream f(fn, std::ios::binary);
    f.write(reinterpret_cast<const char*>(&rows), sizeof(int));
    f.write(reinterpret_cast<const char*>(&cols), sizeof(int));
    f.write(reinterpret_cast<const char*>(m), sizeof(double) * (size_t)rows * cols);
}

struct Handle {
    xm_handle* h = nullptr;
    Handle() {
        xm_options o;
        xm_default_options(&o);
        o.verbose = 1;
        int rc = xm_create(&h, &o);
        if (rc != XM_OK) throw std::runtime_error("xm_create failed (code " + std::to_string(rc) + "): an sm_100 GPU is required, there is no CPU fallback");
    }
    ~Handle() { xm_destroy(h); }
    void check(int rc, const char* what) {
        if (rc != XM_OK) throw std::runtime_error(std::string(what) + " failed: " + xm_last_error(h));
    }
};

void banner(const char* text) {
    std::cout << "+++++++++++++++++++++++++++++++++" << std::endl;
    std::cout << text << std::endl;
    std::cout << "+++++++++++++++++++++++++++++++++" << std::endl;
}

enum class Mode { Full, Rank3, Rebuttle };

// The rank staircase of XM_main.cu:180-310 (solve), :312-401 (solve_rank3) and :35-178 (solve_rebuttle).
int run(const std::string& dataset_path, unsigned max_rank, double tol, double lam, double max_time, Mode mode) {
    py::gil_scoped_release nogil;     // the reference holds the GIL for the whole call; nothing here needs it
    banner("Begin XM");
    const std::string out = dataset_path + "/";
    std::vector<double> Q;
    int rows = 0, cols = 0;
    if (!load_bin(dataset_path + "/Q.bin", Q, rows, cols) || rows <= 0 || rows != cols || rows % 3)
        throw std::runtime_error("Q.bin missing or not a square 3N x 3N matrix");
    const int n3 = rows, n = rows / 3;
    Handle H;
    H.check(xm_set_q_dense(H.h, n3, Q.data(), n3), "xm_set_q_dense");
    std::vector<double>().swap(Q);

    std::vector<double> s_ini;
    if (mode == Mode::Rebuttle) {     // XM_main.cu:61-63: loaded, then R0 is overwritten by the identity at o == 3 (:95-103)
        std::vector<double> t; int rr, cc;
        load_bin(dataset_path + "/R_ini.bin", t, rr, cc);
        if (load_bin(dataset_path + "/s_ini.bin", t, rr, cc) && (int)t.size() == n) s_ini = t;
    }
    // the staircase itself lives behind the C-ABI (xm_solve): the same code serves block-CSR operators and communicators
    const int cap = (int)std::max(3u, std::min(max_rank, 20u));
    std::vector<double> R0((size_t)n3 * cap), s0(n);
    xm_solve_result res;
    const int cmode = mode == Mode::Full ? XM_MODE_FULL : mode == Mode::Rank3 ? XM_MODE_RANK3 : XM_MODE_REBUTTLE;
    H.check(xm_solve(H.h, cmode, (int)max_rank, tol, lam, max_time, s_ini.empty() ? nullptr : s_ini.data(), XM_CERT_AUTO,
                     R0.data(), s0.data(), &res), "xm_solve");
    const int status = res.status;
    const unsigned o = (unsigned)res.rank + 1;
    save_bin(out + "R.bin", R0.data(), n3, (int)(o - 1));
    std::cout << "saved R" << std::endl;
    save_bin(out + "s.bin", s0.data(), n, 1);
    return status;
}

void solve(const std::string& p, unsigned max_rank, double tol, double lam, double max_time) { run(p, max_rank, tol, lam, max_time, Mode::Full); }
void solve_rank3(const std::string& p, unsigned max_rank, double tol, double lam, double max_time) { run(p, max_rank, tol, lam, max_time, Mode::Rank3); }
int solve_rebuttle(const std::string& p, unsigned max_rank, double tol, double lam, double max_time) { return run(p, max_rank, tol, lam, max_time, Mode::Rebuttle); }

}  // namespace

PYBIND11_MODULE(XM, m) {
    m.doc() = "pybind11 for XM (B200-native build: libxm_b200.so behind the reference's Python surface)";
    m.def("solve", &solve, "XM main function");
    m.def("solve_rebuttle", &solve_rebuttle, "permit give initial guess");
    m.def("solve_rank3", &solve_rank3, "XM main function for rank 3 only");
}
