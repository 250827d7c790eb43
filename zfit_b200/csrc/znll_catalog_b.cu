// Ahead-of-time kernel catalogue, part B: signal + background mass fits (sm_100a).
#include "znll_catalog.cuh"
using namespace znll;

static const ZnllCatalogEntry kEntries[] = {
    ZNLL_ENTRY(Sum<CBall<0>,Expo<0>>),
    ZNLL_ENTRY(Sum<CBall<0>,Cheb<0,1>>),
    ZNLL_ENTRY(Sum<CBall<0>,Cheb<0,2>>),
    ZNLL_ENTRY(Sum<Gauss<0>,Cheb<0,1>>),
    ZNLL_ENTRY(Sum<Gauss<0>,Cheb<0,2>>),
    ZNLL_ENTRY(Sum<Gauss<0>,Cheb<0,3>>),
    ZNLL_ENTRY(Sum<Gauss<0>,Cheb<0,4>>),
    ZNLL_ENTRY(Sum<Gauss<0>,Gauss<0>,Expo<0>>),
    ZNLL_ENTRY(Sum<CBall<0>,Gauss<0>,Expo<0>>),
};
static Registrar reg_b(kEntries, sizeof(kEntries) / sizeof(kEntries[0]));
