/* Material model descriptor of the host layer (reference API: solver-large/fea_model.h).
 * The reference dispatches stress / tangent through function pointers; function pointers
 * cannot cross to the device, so here the descriptor only carries the model id and its
 * parameters and the CUDA kernels switch on the id. */
#ifndef FEA_B200_FEA_MODEL_H
#define FEA_B200_FEA_MODEL_H
#include "dense_matrix.h"

typedef enum { MODEL_A5, MODEL_COMPRESSIBLE_NEOHOOKEAN } model_type;

typedef struct fea_model {
  model_type model;
  real parameters[MAX_MATERIAL_PARAMETERS]; /* [0] = lambda, [1] = mu (fea_model.c:38-39) */
  int parameters_count;
} fea_model;
typedef fea_model *fea_model_ptr;

/* validates the id (the reference installs its CPU function pointers here, fea_model.c:7) */
void fea_model_init(fea_model_ptr self, model_type type);
const char *fea_model_name(model_type type);

#endif
