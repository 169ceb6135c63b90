// quartet classes with bra pair class 3 (l_a=2, l_b=0); see eri_class_inst.cuh
#define UNOMOL_BRA_LA 2
#define UNOMOL_BRA_LB 0
#include "eri_class_inst.cuh"
