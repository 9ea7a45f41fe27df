// Host-side checker for the device arithmetic headers (fp256.cuh / curve.cuh).
// Compiled with plain g++ (tests/test_host_arith.py); the ptx_arith.cuh wrappers fall back to their carry-flag
// emulation, so the very same limb algorithms the GPU runs are executed here and compared with the Python oracle.
// Test scaffolding only.
//
// stdin: one command per line, operands as hex integers (canonical, non-Montgomery).  stdout: hex results.
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../octopuszk_b200/csrc/curve.cuh"

using namespace ozk;

template <class F>
static F parse(const std::string& s) {
    F r = F::zero();
    std::string t = std::string(64 - s.size(), '0') + s;
    for (int i = 0; i < 8; i++) r.v[7 - i] = (uint32_t)strtoul(t.substr(i * 8, 8).c_str(), nullptr, 16);
    return r;
}
template <class F>
static std::string hex(const F& a) {
    char buf[80];
    std::string o;
    for (int i = 7; i >= 0; i--) {
        snprintf(buf, sizeof buf, "%08x", a.v[i]);
        o += buf;
    }
    return o;
}
static Fq2 parse2(const std::string& a, const std::string& b) { return {Fq::to_mont(parse<Fq>(a)), Fq::to_mont(parse<Fq>(b))}; }
static std::string hex2(const Fq2& a) { return hex(Fq::from_mont(a.c0)) + " " + hex(Fq::from_mont(a.c1)); }

template <class F>
static void field_cmd(const std::string& op, std::istringstream& in) {
    std::string sa, sb;
    in >> sa;
    F a = F::to_mont(parse<F>(sa));
    F r;
    if (op == "neg") r = F::neg(a);
    else if (op == "inv") r = F::inv(a);
    else if (op == "sqr") r = F::sqr(a);
    else if (op == "dbl") r = F::dbl(a);
    else if (op == "canon") { std::cout << (parse<F>(sa).is_canonical() ? 1 : 0) << "\n"; return; }
    else if (op == "lz") {
        // lazy-domain butterfly on RAW 256-bit values a, b in [0, 2m) and a canonical twiddle w: prints the raw results of
        // add_lazy, sub_lazy, mul_lazy(sub_lazy_wide(a, b), w R) and reduce_lazy(add_lazy) -- the test checks ranges and residues
        std::string sw;
        in >> sb >> sw;
        F x = parse<F>(sa), y = parse<F>(sb), w = F::to_mont(parse<F>(sw));
        F s = F::add_lazy(x, y);
        std::cout << hex(s) << " " << hex(F::sub_lazy(x, y)) << " " << hex(F::mul_lazy(F::sub_lazy_wide(x, y), w)) << " "
                  << hex(F::reduce_lazy(s)) << "\n";
        return;
    }
    else {
        in >> sb;
        F b = F::to_mont(parse<F>(sb));
        if (op == "mul") r = F::mul(a, b);
        else if (op == "lazymul") r = F::redc(F::mul_wide(a, b));
        else if (op == "lazydiff") {
            // a*b - b*b via one reduction
            r = F::redc(F::wide_sub_lazy(F::mul_wide(a, b), F::mul_wide(b, b)));
        } else if (op == "lazysum") {
            // redc of (a*b + a*a - a*b) exercising wide_add / wide_sub
            r = F::redc(F::wide_sub(F::wide_add(F::mul_wide(a, b), F::mul_wide(a, a)), F::mul_wide(a, b)));
        }
        else if (op == "add") r = F::add(a, b);
        else if (op == "sub") r = F::sub(a, b);
        else { std::cout << "bad\n"; return; }
    }
    std::cout << hex(F::from_mont(r)) << "\n";
}

// G1 points: affine x y (0 0 = infinity)
static G1Affine g1_parse(std::istringstream& in) {
    std::string x, y;
    in >> x >> y;
    return {Fq::to_mont(parse<Fq>(x)), Fq::to_mont(parse<Fq>(y))};
}
static void g1_print(const G1XYZZ& p) {
    G1Affine a = xyzz_to_affine(p);
    G1Jac j = xyzz_to_jacobian(p);
    // also check the inversion-free Jacobian conversion: x = X/Z^2, y = Y/Z^3
    if (!p.is_inf()) {
        Fq zi = Fq::inv(j.z), zi2 = Fq::sqr(zi);
        Fq x = Fq::mul(j.x, zi2), y = Fq::mul(j.y, Fq::mul(zi2, zi));
        if (x != a.x || y != a.y) { std::cout << "jacobian-mismatch\n"; return; }
    }
    std::cout << hex(Fq::from_mont(a.x)) << " " << hex(Fq::from_mont(a.y)) << "\n";
}
static G2Affine g2_parse(std::istringstream& in) {
    std::string a, b, c, d;
    in >> a >> b >> c >> d;
    return {parse2(a, b), parse2(c, d)};
}
static void g2_print(const G2XYZZ& p) {
    G2Affine a = xyzz_to_affine(p);
    G2Jac j = xyzz_to_jacobian(p);
    if (!p.is_inf()) {
        Fq2 zi = Fq2::inv(j.z), zi2 = Fq2::sqr(zi);
        Fq2 x = Fq2::mul(j.x, zi2), y = Fq2::mul(j.y, Fq2::mul(zi2, zi));
        if (x != a.x || y != a.y) { std::cout << "jacobian-mismatch\n"; return; }
    }
    std::cout << hex2(a.x) << " " << hex2(a.y) << "\n";
}

int main() {
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream in(line);
        std::string fld, op;
        in >> fld >> op;
        if (fld == "fq") field_cmd<Fq>(op, in);
        else if (fld == "fr") field_cmd<Fr>(op, in);
        else if (fld == "fq2") {
            std::string a, b, c, d;
            in >> a >> b;
            Fq2 x = parse2(a, b), r;
            if (op == "sqr") r = Fq2::sqr(x);
            else if (op == "inv") r = Fq2::inv(x);
            else {
                in >> c >> d;
                Fq2 y = parse2(c, d);
                r = op == "mul" ? Fq2::mul(x, y) : op == "add" ? Fq2::add(x, y) : Fq2::sub(x, y);
            }
            std::cout << hex2(r) << "\n";
        } else if (fld == "g1") {
            // g1 madd_chain k  p0 p1 ... : XYZZ accumulator += each affine point in turn
            // g1 add  a b c d : (a madd b) xyzz_add (c madd d)
            if (op == "chain") {
                int k; in >> k;
                G1XYZZ acc = G1XYZZ::inf();
                for (int i = 0; i < k; i++) { G1Affine p = g1_parse(in); xyzz_madd(acc, p); }
                g1_print(acc);
            } else if (op == "add") {
                G1XYZZ u = G1XYZZ::inf(), w = G1XYZZ::inf();
                G1Affine a = g1_parse(in), b = g1_parse(in), c = g1_parse(in), d = g1_parse(in);
                xyzz_madd(u, a); xyzz_madd(u, b); xyzz_madd(w, c); xyzz_madd(w, d);
                xyzz_add(u, w);
                g1_print(u);
            } else if (op == "dbl") {
                G1XYZZ u = G1XYZZ::inf();
                G1Affine a = g1_parse(in), b = g1_parse(in);
                xyzz_madd(u, a); xyzz_madd(u, b);
                g1_print(xyzz_dbl(u));
            } else if (op == "coopdbl") {
                // k doublings with the level schedule the MSM tail runs across the lanes of a warp (here: a plain loop per level)
                int k; in >> k;
                G1XYZZ u = G1XYZZ::inf();
                G1Affine a = g1_parse(in), b = g1_parse(in);
                xyzz_madd(u, a); xyzz_madd(u, b);
                for (int i = 0; i < k; i++) u = xyzz_dbl_levels(u, SerialMul4<Fq>());
                g1_print(u);
            } else if (op == "coopadd") {
                G1XYZZ u = G1XYZZ::inf(), w = G1XYZZ::inf();
                G1Affine a = g1_parse(in), b = g1_parse(in), c = g1_parse(in), d = g1_parse(in);
                xyzz_madd(u, a); xyzz_madd(u, b); xyzz_madd(w, c); xyzz_madd(w, d);
                g1_print(xyzz_add_levels(u, w, SerialMul4<Fq>()));
            }
        } else if (fld == "g2") {
            if (op == "chain") {
                int k; in >> k;
                G2XYZZ acc = G2XYZZ::inf();
                for (int i = 0; i < k; i++) { G2Affine p = g2_parse(in); xyzz_madd(acc, p); }
                g2_print(acc);
            } else if (op == "add") {
                G2XYZZ u = G2XYZZ::inf(), w = G2XYZZ::inf();
                G2Affine a = g2_parse(in), b = g2_parse(in), c = g2_parse(in), d = g2_parse(in);
                xyzz_madd(u, a); xyzz_madd(u, b); xyzz_madd(w, c); xyzz_madd(w, d);
                xyzz_add(u, w);
                g2_print(u);
            } else if (op == "dbl") {
                G2XYZZ u = G2XYZZ::inf();
                G2Affine a = g2_parse(in), b = g2_parse(in);
                xyzz_madd(u, a); xyzz_madd(u, b);
                g2_print(xyzz_dbl(u));
            } else if (op == "coopdbl") {
                // k doublings with the level schedule the MSM tail runs across the lanes of a warp (here: a plain loop per level)
                int k; in >> k;
                G2XYZZ u = G2XYZZ::inf();
                G2Affine a = g2_parse(in), b = g2_parse(in);
                xyzz_madd(u, a); xyzz_madd(u, b);
                for (int i = 0; i < k; i++) u = xyzz_dbl_levels(u, SerialMul4<Fq2>());
                g2_print(u);
            } else if (op == "coopadd") {
                G2XYZZ u = G2XYZZ::inf(), w = G2XYZZ::inf();
                G2Affine a = g2_parse(in), b = g2_parse(in), c = g2_parse(in), d = g2_parse(in);
                xyzz_madd(u, a); xyzz_madd(u, b); xyzz_madd(w, c); xyzz_madd(w, d);
                g2_print(xyzz_add_levels(u, w, SerialMul4<Fq2>()));
            }
        } else {
            std::cout << "bad\n";
        }
    }
    return 0;
}
