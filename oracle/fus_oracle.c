/*
 * ORACLE - TEST INFRASTRUCTURE ONLY (see fus_oracle_impl.h).
 * CPU restatement of fenicsx-fus-gpu's matrix-free wave hot path in plain C.
 * Built by oracle/Makefile into oracle/libfus_oracle.so; loaded by
 * oracle/oracle.py.  Importers allowed: tests/, __graft_entry__.smoke(),
 * bench.py's cpu_baseline / --impl reference legs.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define T float
#define SFX f32
#include "fus_oracle_impl.h"
#undef T
#undef SFX

#define T double
#define SFX f64
#include "fus_oracle_impl.h"
#undef T
#undef SFX
