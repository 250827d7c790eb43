// Ahead-of-time kernel catalogue, part A: single leaves and 1-D sums (sm_100a).
#include "znll_catalog.cuh"
using namespace znll;

static const ZnllCatalogEntry kEntries[] = {
    ZNLL_ENTRY(Gauss<0>),
    ZNLL_ENTRY(Expo<0>),
    ZNLL_ENTRY(CBall<0>),
    ZNLL_ENTRY(Cheb<0,1>),
    ZNLL_ENTRY(Cheb<0,2>),
    ZNLL_ENTRY(Cheb<0,3>),
    ZNLL_ENTRY(Cheb<0,4>),
    ZNLL_ENTRY(Sum<Gauss<0>,Expo<0>>),
    ZNLL_ENTRY(Sum<Gauss<0>,Gauss<0>>),
    ZNLL_ENTRY(Sum<Gauss<0>,Gauss<0>,Gauss<0>>),
};
static Registrar reg_a(kEntries, sizeof(kEntries) / sizeof(kEntries[0]));
