// oracle/shim: stand-in for the absent aicodix/code header of this name (see shim_code.hh)
#pragma once
#include "shim_code.hh"
