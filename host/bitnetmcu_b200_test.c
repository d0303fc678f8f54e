/*
 * bitnetmcu_b200_test.c -- plain-C mirror of /root/reference/BitNetMCU_MNIST_test.c:17-40: run the test digits of a
 * BitNetMCU_MNIST_test_data.h through the engine and print the reference's own "label: %d predicted: %d" lines
 * (BASELINE.json config #1 plumbing check: stdout must equal the reference's, its exit code is garbage by design).
 * Build: cc -Iinclude -I<dir with BitNetMCU_model.h and BitNetMCU_MNIST_test_data.h> host/bitnetmcu_b200_test.c \
 *           host/bitnetmcu_b200_dll.c -Lbitnetmcu_b200 -lbitnetmcu_b200 -Wl,-rpath,<abs path of bitnetmcu_b200>
 */
#include <stdint.h>
#include <stdio.h>

#include "BitNetMCU_MNIST_test_data.h"

uint32_t BitMnistInference(int8_t *);

int main(void)
{
    int8_t *inputs[] = {input_data_0, input_data_1, input_data_2, input_data_3, input_data_4,
                        input_data_5, input_data_6, input_data_7, input_data_8, input_data_9};
    uint8_t labels[] = {label_0, label_1, label_2, label_3, label_4, label_5, label_6, label_7, label_8, label_9};
    for (int i = 0; i < 10; i++) {
        uint8_t predicted_label = (uint8_t)BitMnistInference(inputs[i]);
        printf("label: %d predicted: %d\n", labels[i], predicted_label);
    }
    return 0;
}
