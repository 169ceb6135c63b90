// quartet classes with bra pair class 1 (l_a=1, l_b=0); see eri_class_inst.cuh
#define UNOMOL_BRA_LA 1
#define UNOMOL_BRA_LB 0
#include "eri_class_inst.cuh"
