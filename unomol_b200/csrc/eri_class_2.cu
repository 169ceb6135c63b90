// quartet classes with bra pair class 2 (l_a=1, l_b=1); see eri_class_inst.cuh
#define UNOMOL_BRA_LA 1
#define UNOMOL_BRA_LB 1
#include "eri_class_inst.cuh"
