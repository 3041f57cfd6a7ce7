// cf_rng.h -- host facades of the two generators (sobol.h:31-151, mrg32k3a.h:23-394).
//
// The streams themselves live on the device (cf_device.cuh: per-thread Gray-code Sobol, integer
// mrg32k3a with matrix-power skip-ahead).  Inside the simulation algorithms only deviceImage() is
// used.  The sequential RNG interface (init / nextU / nextG / skipTo, mcBase.h:228-246) is still
// functional: it serves blocks of paths drawn by the device kernels (cf_rng_draw), so that the
// numbers are the very ones the path kernels consume.
#pragma once

#include "cf_base.h"

class DeviceBackedRNG : public RNG
{
protected:
    size_t              myDim = 0;
    unsigned            myIndex = 0;       // next path to serve
    std::vector<double> myBlockU, myBlockG;
    unsigned            myBlockFirst = 0;
    size_t              myBlockCount = 0;
    bool                myHaveU = false, myHaveG = false;
    static constexpr size_t BLOCK = 1024;

    void refill(const bool gaussian)
    {
        cf_rng img{};
        deviceImage(img);
        std::vector<double>& blk = gaussian ? myBlockG : myBlockU;
        blk.resize(BLOCK * myDim);
        cfCheck(cf_rng_draw(&img, int(myDim), myIndex, BLOCK, gaussian ? 1 : 0, blk.data()));
        if (myBlockFirst != myIndex || myBlockCount == 0) { myHaveU = myHaveG = false; }
        myBlockFirst = myIndex;
        myBlockCount = BLOCK;
        (gaussian ? myHaveG : myHaveU) = true;
    }
    void serve(std::vector<double>& out, const bool gaussian)
    {
        const bool have = gaussian ? myHaveG : myHaveU;
        if (!have || myIndex < myBlockFirst || myIndex >= myBlockFirst + myBlockCount) refill(gaussian);
        const std::vector<double>& blk = gaussian ? myBlockG : myBlockU;
        std::copy(blk.begin() + size_t(myIndex - myBlockFirst) * myDim,
                  blk.begin() + size_t(myIndex - myBlockFirst + 1) * myDim, out.begin());
        ++myIndex;
    }

public:
    void init(const size_t simDim) override { myDim = simDim; myIndex = 0; myBlockCount = 0; myHaveU = myHaveG = false; }
    void nextU(std::vector<double>& uVec) override { serve(uVec, false); }
    void nextG(std::vector<double>& gaussVec) override { serve(gaussVec, true); }
    void skipTo(const unsigned b) override { myIndex = b; }
};

// Sobol's sequence, Joe-Kuo "old 1111" direction numbers, Gray-code order (sobol.h:31-151)
class Sobol : public DeviceBackedRNG
{
public:
    std::unique_ptr<RNG> clone() const override { return std::make_unique<Sobol>(*this); }
    bool deviceImage(cf_rng& img) const override { img.kind = CF_RNG_SOBOL; img.seed1 = img.seed2 = 0; return true; }
    // the reference's skipTo(0) is a no-op (sobol.h:122); path indices are absolute here as well
};

// L'Ecuyer's MRG32k3a with antithetic pairing (mrg32k3a.h:23-394)
class mrg32k3a : public DeviceBackedRNG
{
    unsigned myA, myB;
public:
    mrg32k3a(const unsigned a = 12345, const unsigned b = 12346) : myA(a), myB(b) {}
    std::unique_ptr<RNG> clone() const override { return std::make_unique<mrg32k3a>(*this); }
    bool deviceImage(cf_rng& img) const override { img.kind = CF_RNG_MRG32K3A; img.seed1 = myA; img.seed2 = myB; return true; }
    unsigned seedA() const { return myA; }
    unsigned seedB() const { return myB; }
};
