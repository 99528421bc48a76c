/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * Drives the UNMODIFIED reference QPS reader (interfaces/qps/src/qpalm_qps.c: get_sizes_and_check_format + read_data,
 * both non-static) the way its own main() does (qpalm_qps.c:692-770) and hands the resulting QPALMData back, so that
 * tests/test_qps.py can compare qpalm_b200_qps_read with the reference on the same file.  Compiled by oracle/Makefile
 * into oracle/_ref/libqpalm_qps_ref.so together with the reference sources where they lie.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "qpalm.h"
#include "index_hash.h"
#include "qps_conversion.h"

typedef struct { int no_name_bounds; int no_name_rhs; } read_options;   /* private typedef of qpalm_qps.c:16-19 */
int get_sizes_and_check_format(FILE *fp, QPALMData *data, struct index_table **free_bounds, struct list *free_bounds_list,
                               read_options *opts);
void read_data(FILE *fp, QPALMData *data, struct index_table *row_index_table, struct index_table *col_index_table,
               struct index_table *free_bounds, struct list *free_bounds_list, read_options *opts);
void read_settings(QPALMSettings *settings, FILE *fp);

QPALMData *qps_ref_read(const char *path) {
  FILE *fp = fopen(path, "r");
  if (!fp) return NULL;
  char line[100], command[20], name[50];
  if (!fgets(line, 100, fp) || sscanf(line, "%s %s", command, name) != 2 || strcmp(command, "NAME")) { fclose(fp); return NULL; }
  char *file_copy = NULL;
  struct index_table *free_bounds = NULL;
  struct list *free_bounds_list = list_create();
  QPALMData *data = calloc(1, sizeof(QPALMData));
  read_options opts = {0, 0};
  if (get_sizes_and_check_format(fp, data, &free_bounds, free_bounds_list, &opts)) {
    file_copy = convert_qps_to_new_format(path);
    fp = fopen(file_copy, "r");
    if (!fp) return NULL;
    fgets(line, 100, fp);
    get_sizes_and_check_format(fp, data, &free_bounds, free_bounds_list, &opts);
  }
  fp = fopen(file_copy ? file_copy : path, "r");
  if (!fp) return NULL;
  size_t m = data->m, n = data->n;
  c_int tn = (c_int)n / 5 > 1 ? (c_int)n / 5 : 1;
  size_t n_bounds = n - length_table(free_bounds, (size_t)tn);
  c_int tm = (c_int)(m - n_bounds) / 5 > 1 ? (c_int)(m - n_bounds) / 5 : 1;
  struct index_table *rows = create_index_table(tm), *cols = create_index_table(tn);
  read_data(fp, data, rows, cols, free_bounds, free_bounds_list, &opts);
  fclose(fp);
  if (file_copy) { remove(file_copy); free(file_copy); }
  return data;
}

void qps_ref_read_settings(const char *path, QPALMSettings *settings) {
  FILE *fp = fopen(path, "r");
  if (!fp) { qpalm_set_default_settings(settings); return; }
  read_settings(settings, fp);
  fclose(fp);
}
