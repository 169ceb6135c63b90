// quartet classes with bra pair class 0 (l_a=0, l_b=0); see eri_class_inst.cuh
#define UNOMOL_BRA_LA 0
#define UNOMOL_BRA_LB 0
#include "eri_class_inst.cuh"
