// oracle/shim: stand-in for the absent aicodix/dsp header of this name (see shim_dsp.hh)
#pragma once
#include "shim_dsp.hh"
