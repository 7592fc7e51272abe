// TEST INFRASTRUCTURE ONLY (see orc_real.hpp).
// Avionics of the C172X variants (control laws + direct guidance). Placeholder: NoAvionics only.
#pragma once
#include "orc_vehicle.hpp"
struct Avionics {
    int variant = VAR_C172S;
    bool set_input(int, double) { return false; }
    double get_input(int) const { return 0.0 / 0.0; }
    double output(const Member&, int) const { return 0.0 / 0.0; }
    void assign(Member&) {}
    void f_step(Member&) {}
    void f_periodic(Member&, const World&, double) {}
    void f_init(Member&, const World&) {}
};
