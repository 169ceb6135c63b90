// quartet classes with bra pair class 4 (l_a=2, l_b=1); see eri_class_inst.cuh
#define UNOMOL_BRA_LA 2
#define UNOMOL_BRA_LB 1
#include "eri_class_inst.cuh"
