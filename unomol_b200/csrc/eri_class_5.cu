// quartet classes with bra pair class 5 (l_a=2, l_b=2); see eri_class_inst.cuh
#define UNOMOL_BRA_LA 2
#define UNOMOL_BRA_LB 2
#include "eri_class_inst.cuh"
