/* constraints.h -- drop-in for plonk.c's src/constraints.h: gates, copy constraints, assignments and the
 * expression-to-gates helper (host-side circuit authoring; src/constraints.h:11-309). */
#ifndef CONSTRAINTS_H
#define CONSTRAINTS_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "hf.h"

/* q_l*a + q_r*b + q_o*c + q_m*a*b + q_c = 0 */
typedef struct {
  HF q_l;
  HF q_r;
  HF q_o;
  HF q_m;
  HF q_c;
} GATE;

typedef enum { COPYOF_A, COPYOF_B, COPYOF_C } COPY_OF_TYPE;

typedef struct {
  COPY_OF_TYPE type;
  size_t index; /* 1-based (src/plonk.h:144) */
} COPY_OF;

typedef struct {
  HF *q_l;
  HF *q_r;
  HF *q_o;
  HF *q_m;
  HF *q_c;
  size_t num_gates;

  COPY_OF *c_a;
  COPY_OF *c_b;
  COPY_OF *c_c;
  size_t num_constraints;
} CONSTRAINTS;

typedef struct {
  HF a;
  HF b;
  HF c;
} ASSIGNMENT;

typedef struct {
  HF *a;
  HF *b;
  HF *c;
  size_t len;
} ASSIGNMENTS;

typedef enum { EXPR_VAR, EXPR_CONST, EXPR_SUM, EXPR_SUB, EXPR_MUL } EXPR_TYPE;

typedef struct expression {
  EXPR_TYPE type;
  union {
    const char *var_name;
    HF const_value;
    struct {
      struct expression *left;
      struct expression *right;
    } binary;
  } data;
} EXPRESSION;

#define MAX_VARS 100

typedef struct {
  char *names[MAX_VARS];
  size_t indices[MAX_VARS];
  size_t count;
} VAR_MAP;

typedef struct {
  GATE *gates;
  size_t *a_indices;
  size_t *b_indices;
  size_t *c_indices;
  size_t num_gates;
  size_t capacity;
} GATE_LIST;

#ifdef __cplusplus
extern "C" {
#endif
GATE gate_new(HF q_l, HF q_r, HF q_o, HF q_m, HF q_c);
GATE gate_sum_a_b(void);
GATE gate_sub_a_b(void);
GATE gate_mul_a_b(void);
GATE gate_bind_a(HF value);
GATE gate_bind_to_zero(void);
CONSTRAINTS constraints_new(GATE *gates, size_t num_gates, COPY_OF *c_a, COPY_OF *c_b, COPY_OF *c_c, size_t num_constraints);
bool constraints_satisfy(const CONSTRAINTS *c, const ASSIGNMENTS *a);
void constraints_free(CONSTRAINTS *cons);

void var_map_init(VAR_MAP *vm);
size_t var_map_get_or_add(VAR_MAP *vm, const char *name);
const char *var_map_get_name(VAR_MAP *vm, size_t index);
void var_map_free(VAR_MAP *vm);
void gate_list_init(GATE_LIST *gl);
void gate_list_append(GATE_LIST *gl, GATE g, size_t a_index, size_t b_index, size_t c_index);
void gate_list_free(GATE_LIST *gl);
size_t eval_expr(EXPRESSION *expr, VAR_MAP *vars, GATE_LIST *gates);
#ifdef __cplusplus
}
#endif

#endif /* CONSTRAINTS_H */
