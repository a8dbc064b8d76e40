// GimicB200Interface -- header-only C++ wrapper with the method set of the reference's GimicInterface
// (src/libgimic/GimicInterface.h:4-15, GimicInterface.cpp:7-41), over the HANDLE-based C ABI of include/gimic_b200.h:
// several objects can coexist (the reference's class drives one hidden Fortran context), errors are C++ exceptions
// instead of a Fortran `stop`, and the batched calls are exposed next to the one-point calls.
//
//     GimicB200Interface g("MOL", "XDENS");            // was: GimicInterface g("MOL", "XDENS");
//     g.set_magnet(b); g.calc_jvector(r, jv);           // unchanged call sites
//     g.calc_jvectors(n, points, jvecs);                // what a grid loop should call instead
//
// The reference's own GimicInterface.cpp also compiles and links unchanged against libgimic_b200.so (legacy symbols).
#ifndef GIMIC_B200_INTERFACE_H
#define GIMIC_B200_INTERFACE_H

#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>

#include "gimic_b200.h"

class GimicB200Interface {
  public:
    // screening_thrs: gimic_init uses SCREEN_THRS = 1e-6 (globals.f90:56, gimic_interface.f90:41); the gimic.inp default is 1e-8
    GimicB200Interface(const char *mol, const char *xdens, int uhf = 0, double screening_thrs = 1.0e-6, int device = -1)
        : h_(0), mol_(mol ? mol : ""), xdens_(xdens ? xdens : ""), spin_(GIMIC_B200_TOTAL) {
        gimic_b200_default_opts(&opts_);
        opts_.uhf = uhf ? 1 : 0; opts_.screening_thrs = screening_thrs; opts_.device = device;
        magnet_[0] = magnet_[1] = magnet_[2] = 0.0;
        create();
    }
    virtual ~GimicB200Interface() { if (h_) gimic_b200_destroy(h_); }

    // ---- the reference's methods (GimicInterface.h:8-14) ----------------------------------------------------------
    void set_uhf(int uhf) {   // the reference only flips a flag and then reads unallocated beta densities; here the context is rebuilt
        if ((uhf != 0) == (opts_.uhf != 0)) return;
        opts_.uhf = uhf ? 1 : 0;
        if (h_) { gimic_b200_destroy(h_); h_ = 0; }
        create();
    }
    void set_magnet(const double b[3]) { magnet_[0] = b[0]; magnet_[1] = b[1]; magnet_[2] = b[2]; }
    void set_spin(const char *s) {
        if (!std::strcmp(s, "alpha")) spin_ = GIMIC_B200_ALPHA;
        else if (!std::strcmp(s, "beta")) spin_ = GIMIC_B200_BETA;
        else if (!std::strcmp(s, "total")) spin_ = GIMIC_B200_TOTAL;
        else if (!std::strcmp(s, "spindens")) spin_ = GIMIC_B200_SPINDENS;
        else throw std::invalid_argument(std::string("Invalid spin case: ") + s);       // gimic_interface.f90:111-112
    }
    void set_screening(double thrs) { opts_.screening_thrs = thrs; }   // like the reference: recorded, radii are fixed at construction
    void calc_jtensor(const double r[3], double jt[9]) { calc_jtensors(1, r, jt); }
    void calc_jvector(const double r[3], double jv[3]) { calc_jvectors(1, r, jv); }
    void calc_modj(const double r[3], double *mj) {                    // reference: STOP 'NOT IMPLEMENTED'; here |J|
        double jv[3];
        calc_jvectors(1, r, jv);
        *mj = std::sqrt(jv[0] * jv[0] + jv[1] * jv[1] + jv[2] * jv[2]);
    }

    // ---- batched calls (r: 3 x n point-major; tens: 9 x n, tens[9 i + m + 3 b] = dJ_m/dB_b; flags: GIMIC_B200_DEVICE_PTR) ----
    void calc_jtensors(long n, const double *r, double *tens, int flags = 0) {
        check(gimic_b200_calc_jtensors(h_, n, r, spin_, tens, flags));
    }
    void calc_jvectors(long n, const double *r, double *jvec, int flags = 0) {
        check(gimic_b200_calc_fields(h_, n, r, magnet_, spin_, 0, jvec, 0, 0, 0, 0, 0.0, flags));
    }
    int nbf() const { return gimic_b200_nbf(h_); }
    int natoms() const { return gimic_b200_natoms(h_); }
    gimic_b200_handle handle() const { return h_; }

  private:
    GimicB200Interface(const GimicB200Interface &);              // one owner per context
    GimicB200Interface &operator=(const GimicB200Interface &);
    void create() { check(gimic_b200_create(&h_, mol_.c_str(), xdens_.c_str(), &opts_)); }
    static void check(int rc) { if (rc < 0) throw std::runtime_error(std::string("gimic_b200: ") + gimic_b200_last_error()); }

    gimic_b200_handle h_;
    gimic_b200_opts opts_;
    std::string mol_, xdens_;
    int spin_;
    double magnet_[3];
};

#endif /* GIMIC_B200_INTERFACE_H */
