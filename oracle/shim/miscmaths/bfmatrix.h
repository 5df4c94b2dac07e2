// TEST INFRASTRUCTURE ONLY — FSL bfmatrix.h stand-in (extended for the meshreg build).
#ifndef ORACLE_SHIM_BFMATRIX_H
#define ORACLE_SHIM_BFMATRIX_H
#include "miscmaths/miscmaths.h"
#endif
