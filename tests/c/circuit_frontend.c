/* tests/c/circuit_frontend.c -- TEST PROGRAM.  The reference's circuit authoring API (EXPRESSION, VAR_MAP, GATE_LIST,
 * eval_expr: src/constraints.h:185-309) used exactly as constraints-test.c uses it, followed by the step the reference
 * leaves to the programmer (plonk-test.c:139-213 writes CONSTRAINTS by hand): pb_circuit_from_gates lowers the gate list
 * to a circuit.  Host code only -- runs without a GPU.  Prints one line per circuit:
 *   <name> gates <n> a <idx..> b <idx..> c <idx..> vars <n_vars> circuit <44 bytes>
 */
#include <stdio.h>
#include "constraints.h"
#include "plonk_b200.h"

static EXPRESSION var(const char *name) { EXPRESSION e; e.type = EXPR_VAR; e.data.var_name = name; return e; }
static EXPRESSION bin(EXPR_TYPE t, EXPRESSION *l, EXPRESSION *r) { EXPRESSION e; e.type = t; e.data.binary.left = l; e.data.binary.right = r; return e; }

static int emit(const char *name, GATE_LIST *gl, VAR_MAP *vm, const size_t *eq, size_t n_eq) {
  uint8_t circuit[PB_CIRCUIT_BYTES];
  int rc = pb_circuit_from_gates((const uint8_t *)gl->gates, gl->a_indices, gl->b_indices, gl->c_indices, gl->num_gates, eq, n_eq, circuit);
  if (rc) { printf("%s error %d %s\n", name, rc, pb_last_error()); return rc; }
  printf("%s gates %zu a", name, gl->num_gates);
  for (size_t i = 0; i < gl->num_gates; i++) printf(" %zu", gl->a_indices[i]);
  printf(" b");
  for (size_t i = 0; i < gl->num_gates; i++) printf(" %zu", gl->b_indices[i]);
  printf(" c");
  for (size_t i = 0; i < gl->num_gates; i++) printf(" %zu", gl->c_indices[i]);
  printf(" vars %zu circuit", vm->count);
  for (int i = 0; i < PB_CIRCUIT_BYTES; i++) printf(" %u", circuit[i]);
  printf("\n");
  return 0;
}

int main(void) {
  EXPRESSION x = var("x"), y = var("y"), z = var("z");
  EXPRESSION xx = bin(EXPR_MUL, &x, &x), yy = bin(EXPR_MUL, &y, &y), zz = bin(EXPR_MUL, &z, &z);
  {
    /* the plonk-test circuit, gate by gate: x*x, y*y, z*z compiled from expressions, then x^2 + y^2 = z^2 as one sum gate
     * whose output wire IS the z*z output */
    VAR_MAP vm; GATE_LIST gl;
    var_map_init(&vm); gate_list_init(&gl);
    size_t o1 = eval_expr(&xx, &vm, &gl), o2 = eval_expr(&yy, &vm, &gl), o3 = eval_expr(&zz, &vm, &gl);
    gate_list_append(&gl, gate_sum_a_b(), o1, o2, o3);
    if (emit("plonk_test", &gl, &vm, NULL, 0)) return 1;
    gate_list_free(&gl); var_map_free(&vm);
  }
  {
    /* the same statement as ONE expression tree plus one more: (x*x + y*y) and z*z, outputs asserted equal */
    VAR_MAP vm; GATE_LIST gl;
    var_map_init(&vm); gate_list_init(&gl);
    EXPRESSION lhs = bin(EXPR_SUM, &xx, &yy);
    size_t eq[2];
    eq[0] = eval_expr(&lhs, &vm, &gl);
    eq[1] = eval_expr(&zz, &vm, &gl);
    if (emit("pythagoras_expr", &gl, &vm, eq, 1)) return 1;
    gate_list_free(&gl); var_map_free(&vm);
  }
  {
    /* five gates do not fit the 4-point domain: refused, with the reason */
    VAR_MAP vm; GATE_LIST gl;
    var_map_init(&vm); gate_list_init(&gl);
    EXPRESSION lhs = bin(EXPR_SUM, &xx, &yy), all = bin(EXPR_SUB, &lhs, &zz);
    eval_expr(&all, &vm, &gl);
    uint8_t circuit[PB_CIRCUIT_BYTES];
    int rc = pb_circuit_from_gates((const uint8_t *)gl.gates, gl.a_indices, gl.b_indices, gl.c_indices, gl.num_gates, NULL, 0, circuit);
    printf("five_gates gates %zu rc %d\n", gl.num_gates, rc);
    gate_list_free(&gl); var_map_free(&vm);
  }
  return 0;
}
