// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Turns the reference's UNMODIFIED src/serial/serial.cpp into a linkable function so tests can obtain the
// reference's own table in-process.  Nothing is copied: the Makefile compiles the file where it lies under
// $(REF)/src with  -Dmain=nw_ref_main -DneedlemanWunsch=nw_ref_serial_impl  (the textually included
// driver.cpp's main() gets renamed), and this shim gives it a C ABI.
// dnaArray comes from the reference's own helper.hpp (src/common/helper.hpp:9-12), included in place.
#include <cstdint>
#include "helper.hpp"

void nw_ref_serial_impl(dnaArray s1, dnaArray s2, int* t);   // src/serial/serial.cpp:4 (renamed by -D)

extern "C" void nw_ref_serial_fill(const int8_t* s1, int n1, const int8_t* s2, int n2, int* t)
{
    dnaArray a, b;
    a.size = n1; a.dna = const_cast<int8_t*>(s1);
    b.size = n2; b.dna = const_cast<int8_t*>(s2);
    nw_ref_serial_impl(a, b, t);
}
